// Optimizer step on a flat fp32 arena (clip_grad_norm_ + SGD nesterov) and layout plumbing (weight packing,
// NCDHW <-> NDHWC).
#include "common.cuh"

namespace mtb {

// ---- sum of squares ----------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  pdl_wait();
  double acc = 0.0;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 * 4; i < n; ++i) acc += (double)g[i] * g[i];
  acc = warp_sum(acc);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(out, t);
  }
}

int sumsq(const float* g, long long n, double* out, cudaStream_t s) {
  MTB_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "sumsq: pointer must be 16-byte aligned");
  const int blocks = (int)max(1LL, min((long long)num_sms() * 4, (n / 4 + 255) / 256));
  launch_pdl(sumsq_kernel, dim3(blocks), dim3(256), (size_t)(0), s, g, n, out);
  return check_launch("sumsq");
}

// ---- clip + SGD(nesterov) ------------------------------------------------------------------------------------------
__global__ void sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n,
                                const double* __restrict__ sumsq, float inv_scale, float max_norm, float lr,
                                float momentum, float wd, int first, const float* __restrict__ dyn_scale) {
  pdl_wait();
  const double ss = *sumsq;
  if (!isfinite(ss)) return;  // GradScaler: skip the step on inf/nan gradients
  if (dyn_scale) inv_scale /= dyn_scale[0];  // GradScaler.unscale_: the loss was multiplied by the current scale
  const float total_norm = (float)sqrt(ss) * inv_scale;
  float clip = max_norm / (total_norm + 1e-6f);
  clip = clip > 1.f ? 1.f : clip;
  const float coef = clip * inv_scale;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(g[i], coef, wd * pi);
    const float b = first ? gi : fmaf(momentum, buf[i], gi);
    buf[i] = b;
    p[i] = pi - lr * fmaf(momentum, b, gi);
  }
}

int sgd_step(float* p, const float* g, float* buf, long long n, const double* sumsq_, float inv_scale, float max_norm,
             float lr, float momentum, float wd, int first, const float* dyn_scale, cudaStream_t s) {
  const int blocks = (int)max(1LL, min((long long)num_sms() * 8, (n + 255) / 256));
  launch_pdl(sgd_step_kernel, dim3(blocks), dim3(256), (size_t)(0), s, p, g, buf, n, sumsq_, inv_scale, max_norm, lr, momentum, wd, first, dyn_scale);
  return check_launch("sgd_step");
}

// ---- GradScaler.update on the device ----------------------------------------------------------------------------------
// state = {scale, growth_tracker, found_inf of this step, number of skipped steps}
__global__ void loss_scale_update_kernel(const double* __restrict__ sumsq, float* __restrict__ state, float growth,
                                         float backoff, int interval) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const bool bad = !isfinite(*sumsq);
  float scale = state[0], tracker = state[1];
  if (bad) {
    scale *= backoff;
    tracker = 0.f;
    state[3] += 1.f;
  } else {
    tracker += 1.f;
    if (tracker >= (float)interval) { scale *= growth; tracker = 0.f; }
  }
  state[0] = scale;
  state[1] = tracker;
  state[2] = bad ? 1.f : 0.f;
}

int loss_scale_update(const double* sumsq_, float* state, float growth, float backoff, int interval, cudaStream_t s) {
  loss_scale_update_kernel<<<1, 32, 0, s>>>(sumsq_, state, growth, backoff, interval);
  return check_launch("loss_scale_update");
}

// ---- weight packing ---------------------------------------------------------------------------------------------------
// W(co, ci, t) = transposed ? w[ci][co][t] : w[co][ci][t];  packed[t][r][c] with (r,c) = swap_io ? (ci,co) : (co,ci)
template <typename WT>
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int ntap, int transposed, int swap_io,
                                    WT* __restrict__ packed, int Cout_p, int Cin_p, int split, int split_p) {
  const int R = swap_io ? Cin_p : Cout_p, S = swap_io ? Cout_p : Cin_p;
  const long long n = (long long)ntap * R * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % S);
    const int r = (int)((i / S) % R);
    const int t = (int)(i / ((long long)S * R));
    const int co = swap_io ? c : r;
    int cip = swap_io ? r : c;
    // packed input channel -> logical input channel (two separately padded halves of a concatenated input)
    int ci;
    if (split > 0) ci = cip < split_p ? (cip < split ? cip : -1) : (cip - split_p + split);
    else ci = cip;
    float v = 0.f;
    if (co < Cout && ci >= 0 && ci < Cin)
      v = transposed ? w[((long long)ci * Cout + co) * ntap + t] : w[((long long)co * Cin + ci) * ntap + t];
    Traits<WT>::st(packed + i, v);
  }
}

int pack_weights(const float* w, int Cout, int Cin, int ntap, int transposed, int swap_io, void* packed, int wdtype,
                 int Cout_p, int Cin_p, int split, int split_p, cudaStream_t s) {
  MTB_REQUIRE(Cout <= Cout_p && (split > 0 ? (split <= split_p && Cin - split <= Cin_p - split_p) : Cin <= Cin_p),
              "pack_weights: padded sizes too small (Cout %d/%d Cin %d/%d split %d/%d)", Cout, Cout_p, Cin, Cin_p, split,
              split_p);
  const long long n = (long long)ntap * Cout_p * Cin_p;
  const int blocks = (int)max(1LL, min((long long)num_sms() * 8, (n + 255) / 256));
  MTB_DISPATCH_DTYPE(wdtype, WT, (pack_weights_kernel<WT><<<blocks, 256, 0, s>>>(
      w, Cout, Cin, ntap, transposed, swap_io, reinterpret_cast<WT*>(packed), Cout_p, Cin_p, split, split_p)));
  return check_launch("pack_weights");
}

// ---- all weights of a network in ONE launch --------------------------------------------------------------------------
// The optimizer step invalidates every packed copy, so each training step used to pay ~60 tiny pack launches with
// strided 4-byte gathers.  Here a block owns one (layer, 16 couts x 16 cins) tile for all taps: the reference-layout
// slab is read with consecutive threads on consecutive addresses ([ci][tap] runs of one cout row), staged in shared
// memory, and written out as full 32-byte rows of both packed layouts.
template <typename WT>
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const mtb200_pack_desc* __restrict__ descs, int n) {
  pdl_wait();
  __shared__ float tile[16 * 16 * 33];
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].blk_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const mtb200_pack_desc d = descs[lo];
  const int local = (int)blockIdx.x - d.blk_begin;
  const int tiles_ci = d.Cin_p / 16;
  const int co0 = (local / tiles_ci) * 16, cip0 = (local % tiles_ci) * 16;
  const int ntap = d.ntap, ntp = d.ntap | 1, row = 16 * ntap, total = 256 * ntap;
  for (int e = threadIdx.x; e < total; e += 256) {
    const int a = e / row, rem = e - a * row;
    const int bq = rem / ntap, t = rem - bq * ntap;
    const int co_l = d.transposed ? bq : a, ci_l = d.transposed ? a : bq;
    const int co = co0 + co_l, cip = cip0 + ci_l;
    int ci;
    if (d.split > 0) ci = cip < d.split_p ? (cip < d.split ? cip : -1) : (cip - d.split_p + d.split);
    else ci = cip;
    float v = 0.f;
    if (co < d.Cout && ci >= 0 && ci < d.Cin)
      v = d.transposed ? d.w[((long long)ci * d.Cout + co) * ntap + t] : d.w[((long long)co * d.Cin + ci) * ntap + t];
    tile[(co_l * 16 + ci_l) * ntp + t] = v;
  }
  __syncthreads();
  if (d.packed) {
    WT* out = reinterpret_cast<WT*>(d.packed);
    for (int e = threadIdx.x; e < total; e += 256) {
      const int ci_l = e & 15, co_l = (e >> 4) & 15, t = e >> 8;
      Traits<WT>::st(out + ((long long)t * d.Cout_p + co0 + co_l) * d.Cin_p + cip0 + ci_l, tile[(co_l * 16 + ci_l) * ntp + t]);
    }
  }
  if (d.packed_swap) {
    WT* out = reinterpret_cast<WT*>(d.packed_swap);
    for (int e = threadIdx.x; e < total; e += 256) {
      const int co_l = e & 15, ci_l = (e >> 4) & 15, t = e >> 8;
      Traits<WT>::st(out + ((long long)t * d.Cin_p + cip0 + ci_l) * d.Cout_p + co0 + co_l, tile[(co_l * 16 + ci_l) * ntp + t]);
    }
  }
}

int pack_weights_batched(const void* descs, int n, int total_blocks, int wdtype, cudaStream_t s) {
  if (n <= 0 || total_blocks <= 0) return MTB200_OK;
  MTB_DISPATCH_DTYPE(wdtype, WT, (launch_pdl(pack_weights_batched_kernel<WT>, dim3(total_blocks), dim3(256), (size_t)(0), s, 
      reinterpret_cast<const mtb200_pack_desc*>(descs), n)));
  return check_launch("pack_weights_batched");
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dw, int Cout, int Cin, int ntap, int transposed, int Cout_p,
                                    int Cin_p, int split, int split_p, float scale, int accumulate,
                                    float* __restrict__ grad) {
  const long long n = (long long)Cout * Cin * ntap;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % ntap);
    const int a = (int)((i / ntap) % (transposed ? Cout : Cin));
    const int b = (int)(i / ((long long)ntap * (transposed ? Cout : Cin)));
    const int co = transposed ? a : b, ci = transposed ? b : a;
    const int cip = (split > 0 && ci >= split) ? ci - split + split_p : ci;
    const float v = scale * dw[((long long)t * Cout_p + co) * Cin_p + cip];
    grad[i] = accumulate ? grad[i] + v : v;
  }
}

int unpack_wgrad(const float* dw, int Cout, int Cin, int ntap, int transposed, int Cout_p, int Cin_p, int split,
                 int split_p, float scale, int accumulate, float* grad, cudaStream_t s) {
  const long long n = (long long)Cout * Cin * ntap;
  const int blocks = (int)max(1LL, min((long long)num_sms() * 8, (n + 255) / 256));
  unpack_wgrad_kernel<<<blocks, 256, 0, s>>>(dw, Cout, Cin, ntap, transposed, Cout_p, Cin_p, split, split_p, scale,
                                             accumulate, grad);
  return check_launch("unpack_wgrad");
}

// ---- every weight gradient of a step in ONE launch ----------------------------------------------------------------------
// A block owns UNPACK_CHUNK consecutive elements of one layer's reference-layout gradient (descriptor found by binary
// search on the block index, as in pack_weights_batched): 160 tiny launches per step become one.
constexpr int UNPACK_CHUNK = 4096;

__global__ void __launch_bounds__(256) unpack_wgrad_batched_kernel(const mtb200_unpack_desc* __restrict__ descs, int n) {
  pdl_wait();
  // the chunk's (outer, inner) channel pairs x taps, staged so that BOTH sides are coalesced: the packed buffer
  // [tap][Cout_p][Cin_p] is read tap by tap along consecutive pairs (= consecutive ci of one cout row; the first version
  // gathered one 4-byte element per 32-byte sector and paid three 64-bit divisions per element), the reference-layout
  // gradient [outer][inner][tap] is then updated along consecutive addresses
  __shared__ float stage[UNPACK_CHUNK + 64];
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].blk_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const mtb200_unpack_desc d = descs[lo];
  const long long total = (long long)d.Cout * d.Cin * d.ntap;
  const long long i0 = (long long)((int)blockIdx.x - d.blk_begin) * UNPACK_CHUNK;
  const long long i1 = min(total, i0 + UNPACK_CHUNK);
  const int inner = d.transposed ? d.Cout : d.Cin;
  const int ntap = d.ntap;
  const long long p0 = i0 / ntap;                         // first (outer, inner) pair of the chunk
  const int npair = (int)((i1 - 1) / ntap - p0) + 1;      // <= UNPACK_CHUNK / ntap + 2
  if (npair * ntap <= UNPACK_CHUNK + 64) {
    const int a0 = (int)(p0 % inner), b0 = (int)(p0 / inner);
    for (int e = threadIdx.x; e < npair * ntap; e += 256) {
      const int t = e / npair, pr = e - t * npair;
      int a = a0 + pr, b = b0;
      while (a >= inner) { a -= inner; ++b; }              // a chunk spans a few outer rows at most (or inner is tiny)
      const int co = d.transposed ? a : b, ci = d.transposed ? b : a;
      const int cip = (d.split > 0 && ci >= d.split) ? ci - d.split + d.split_p : ci;
      stage[pr * ntap + t] = b < (d.transposed ? d.Cin : d.Cout) ? d.dw[((long long)t * d.Cout_p + co) * d.Cin_p + cip] : 0.f;
    }
    __syncthreads();
    const long long base = p0 * ntap;
    for (long long i = i0 + threadIdx.x; i < i1; i += 256) d.grad[i] += stage[(int)(i - base)];
    return;
  }
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const int t = (int)(i % d.ntap);
    const int a = (int)((i / d.ntap) % inner);
    const int b = (int)(i / ((long long)d.ntap * inner));
    const int co = d.transposed ? a : b, ci = d.transposed ? b : a;
    const int cip = (d.split > 0 && ci >= d.split) ? ci - d.split + d.split_p : ci;
    d.grad[i] += d.dw[((long long)t * d.Cout_p + co) * d.Cin_p + cip];
  }
}

int unpack_wgrad_batched(const void* descs, int n, int total_blocks, cudaStream_t s) {
  if (n <= 0 || total_blocks <= 0) return MTB200_OK;
  launch_pdl(unpack_wgrad_batched_kernel, dim3(total_blocks), dim3(256), (size_t)(0), s, reinterpret_cast<const mtb200_unpack_desc*>(descs), n);
  return check_launch("unpack_wgrad_batched");
}

// ---- NCDHW fp32 <-> NDHWC --------------------------------------------------------------------------------------------
template <typename T>
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ src, int C, long long nvox, T* __restrict__ dst, int ldc,
                                      int coff, int Cp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long long v0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long v = v0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && v < nvox) ? src[((long long)b * C + c) * nvox + v] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long v = v0 + i;
    const int c = c0 + threadIdx.x;
    if (v < nvox && c < Cp) Traits<T>::st(dst + ((long long)b * nvox + v) * ldc + coff + c, tile[threadIdx.x][i]);
  }
}

// few input channels (the network input: C = 1 padded to 16): one thread per voxel, coalesced plane reads, one 16-byte
// vector store per 8 output channels -- a pure streaming pass instead of 32x32 transposes of mostly padding
template <typename T, int CMAX>
__global__ void ncdhw_to_ndhwc_small_kernel(const float* __restrict__ src, int C, long long nvox, T* __restrict__ dst,
                                            int ldc, int coff, int Cp) {
  const int b = blockIdx.y;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
    float x[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) x[c] = c < C ? src[((long long)b * C + c) * nvox + v] : 0.f;
    T* o = dst + ((long long)b * nvox + v) * ldc + coff;
    for (int c0 = 0; c0 < Cp; c0 += 8) {
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = 0.f;
      if (c0 < CMAX) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < CMAX) y[j] = x[(c0 + j) % CMAX];
      }
      store8<T>(o + c0, y);
    }
  }
}

int ncdhw_to_ndhwc(const float* src, int B, int C, long long nvox, void* dst, int dtype, int ldc, int coff, int Cp,
                   cudaStream_t s) {
  MTB_REQUIRE(C <= Cp && coff + Cp <= ldc, "ncdhw_to_ndhwc: C=%d Cp=%d coff=%d ldc=%d", C, Cp, coff, ldc);
  if (nvox == 0 || B == 0) return MTB200_OK;
  if (C <= 8 && Cp % 8 == 0 && ldc % 8 == 0 && coff % 8 == 0) {
    dim3 grid((unsigned)min((nvox + 255) / 256, (long long)num_sms() * 16), B);
    MTB_DISPATCH_DTYPE(dtype, T, (ncdhw_to_ndhwc_small_kernel<T, 8><<<grid, 256, 0, s>>>(
        src, C, nvox, reinterpret_cast<T*>(dst), ldc, coff, Cp)));
    return check_launch("ncdhw_to_ndhwc");
  }
  dim3 grid((unsigned)((nvox + 31) / 32), (Cp + 31) / 32, B), block(32, 8);
  MTB_DISPATCH_DTYPE(dtype, T, (ncdhw_to_ndhwc_kernel<T><<<grid, block, 0, s>>>(src, C, nvox, reinterpret_cast<T*>(dst),
                                                                                 ldc, coff, Cp)));
  return check_launch("ncdhw_to_ndhwc");
}

template <typename T>
__global__ void ndhwc_to_ncdhw_kernel(const T* __restrict__ src, int ldc, int coff, int C, long long nvox,
                                      float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long long v0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long v = v0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (v < nvox && c < C) ? Traits<T>::ld(src + ((long long)b * nvox + v) * ldc + coff + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long v = v0 + threadIdx.x;
    if (c < C && v < nvox) dst[((long long)b * C + c) * nvox + v] = tile[threadIdx.x][i];
  }
}

int ndhwc_to_ncdhw(const void* src, int dtype, int ldc, int coff, int B, int C, long long nvox, float* dst,
                   cudaStream_t s) {
  if (nvox == 0 || B == 0) return MTB200_OK;
  dim3 grid((unsigned)((nvox + 31) / 32), (C + 31) / 32, B), block(32, 8);
  MTB_DISPATCH_DTYPE(dtype, T, (ndhwc_to_ncdhw_kernel<T><<<grid, block, 0, s>>>(reinterpret_cast<const T*>(src), ldc,
                                                                                 coff, C, nvox, dst)));
  return check_launch("ndhwc_to_ncdhw");
}

}  // namespace mtb
