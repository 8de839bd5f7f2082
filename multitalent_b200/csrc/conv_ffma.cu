// CUDA-core (FFMA, fp32 accumulate) implementation of the tap-table convolution and its weight gradient.
// This is the T0 "parity mode" arithmetic (fp32 storage) and the reference implementation the tcgen05 kernels are
// checked against on the GPU.  See include/mtb200.h for the problem statement.
#include "common.cuh"

namespace mtb {

constexpr int FF_BM = 128;  // output voxels per CTA
constexpr int FF_KC = 16;   // input channels per k-step
constexpr int FF_THREADS = 256;

struct RowInfo { int b, d, h, w; };

__device__ __forceinline__ RowInfo decode_row(long long m, int Do, int Ho, int Wo) {
  RowInfo r;
  r.w = (int)(m % Wo); m /= Wo;
  r.h = (int)(m % Ho); m /= Ho;
  r.d = (int)(m % Do);
  r.b = (int)(m / Do);
  return r;
}

template <typename T, typename WT, int BN>
__global__ void __launch_bounds__(FF_THREADS) conv_taps_ffma_kernel(const __grid_constant__ mtb200_conv_params p) {
  constexpr int TN = BN / 16;  // couts per thread (4 for BN=64, 2 for BN=32)
  __shared__ __align__(16) float As[FF_KC][FF_BM];
  __shared__ __align__(16) float Ws[FF_KC][BN];
  __shared__ RowInfo rows[FF_BM];
  // InstanceNorm statistics: per-thread partial sums over its 8 rows, then a FIXED-ORDER reduction over the 16 row
  // groups (no shared-memory float atomics): the fp32 parity mode is reproducible run to run, so no LeakyReLU branch of a
  // near-zero voxel flips between repetitions (DESIGN.md "Run-to-run variation")
  __shared__ float s_part[2][2][16][BN];  // [sum | sum of squares][sample of the tile: first / second][row group][cout]

  const int tid = threadIdx.x;
  const int g = blockIdx.z;
  const long long M = (long long)p.B * p.Do * p.Ho * p.Wo;
  const long long m0 = (long long)blockIdx.x * FF_BM;
  const int n0 = blockIdx.y * BN;
  const int tap_begin = p.group_tap_begin[g], tap_end = p.group_tap_begin[g + 1];
  const int nkc = p.Cin / FF_KC;
  const int niter = (tap_end - tap_begin) * nkc;

  if (tid < FF_BM) {
    long long m = m0 + tid;
    RowInfo r;
    if (m < M) r = decode_row(m, p.Do, p.Ho, p.Wo); else { r.b = -1; r.d = r.h = r.w = 0; }
    rows[tid] = r;
  }
  __syncthreads();

  const T* __restrict__ in = reinterpret_cast<const T*>(p.in);
  const WT* __restrict__ w = reinterpret_cast<const WT*>(p.w);

  // loader roles
  const int a_row = tid % FF_BM, a_half = tid / FF_BM;            // 8 channels of one voxel row
  const int w_co = tid / 4, w_q = tid % 4;                        // 4 channels of one cout (only tid < 4*BN active)
  const RowInfo ar = rows[a_row];

  float a_reg[8];
  float w_reg[4];

  auto load_regs = [&](int it) {
    const int t = tap_begin + it / nkc;
    const int kc = it % nkc;
    // ---- A
    const int id = ar.d * p.is[0] + p.tap_off[t][0];
    const int ih = ar.h * p.is[1] + p.tap_off[t][1];
    const int iw = ar.w * p.is[2] + p.tap_off[t][2];
    const bool ok = ar.b >= 0 && (unsigned)id < (unsigned)p.Di && (unsigned)ih < (unsigned)p.Hi &&
                    (unsigned)iw < (unsigned)p.Wi;
    if (ok) {
      const int c = kc * FF_KC + a_half * 8;
      const long long vox = (((long long)ar.b * p.Di + id) * p.Hi + ih) * p.Wi + iw;
      load8<T>(in + vox * p.in_ldc + p.in_coff + c, a_reg);
      if (p.xform) {
        const float4* xf = reinterpret_cast<const float4*>(p.xform) + (long long)ar.b * p.Cin + c;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 f = __ldg(xf + j);
          const float v = fmaf(a_reg[j], f.x, f.y);
          a_reg[j] = v > 0.f ? v : v * f.z;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) a_reg[j] = 0.f;
    }
    // ---- W
    if (tid < 4 * BN) {
      const int co = n0 + w_co;
      if (co < p.Cout) {
        const WT* wp = w + ((long long)p.tap_widx[t] * p.Cout + co) * p.Cin + kc * FF_KC + w_q * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) w_reg[j] = Traits<WT>::ld(wp + j);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) w_reg[j] = 0.f;
      }
    }
  };
  auto store_smem = [&]() {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[a_half * 8 + j][a_row] = a_reg[j];
    if (tid < 4 * BN) {
#pragma unroll
      for (int j = 0; j < 4; ++j) Ws[w_q * 4 + j][w_co] = w_reg[j];
    }
  };

  const int tn = tid % 16, tm = tid / 16;  // thread tile: rows tm*8..+8, cols tn*TN..+TN
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (niter > 0) load_regs(0);
  for (int it = 0; it < niter; ++it) {
    store_smem();
    __syncthreads();
    if (it + 1 < niter) load_regs(it + 1);
#pragma unroll
    for (int k = 0; k < FF_KC; ++k) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][tm * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Ws[k][tn * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: bias, round, (accumulate), store, statistics
  T* __restrict__ out = reinterpret_cast<T*>(p.out);
  const int b_first = rows[0].b;
  float bsum[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int co = n0 + tn * TN + j;
    bsum[j] = (p.bias && co < p.Cout) ? p.bias[co] : 0.f;
  }
  float ps0[TN], ps1[TN], pq0[TN], pq1[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) { ps0[j] = ps1[j] = pq0[j] = pq1[j] = 0.f; }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const RowInfo r = rows[tm * 8 + i];
    if (r.b < 0) continue;
    const long long ovox = (((long long)r.b * p.Dof + (r.d * p.os[0] + p.group_ooff[g][0])) * p.Hof +
                            (r.h * p.os[1] + p.group_ooff[g][1])) * p.Wof + (r.w * p.os[2] + p.group_ooff[g][2]);
    T* op = out + ovox * p.out_ldc + p.out_coff + n0 + tn * TN;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + tn * TN + j;
      if (co >= p.Cout) continue;
      float v = acc[i][j] + bsum[j];
      if (p.accumulate) v += Traits<T>::ld(op + j);
      Traits<T>::st(op + j, v);
      if (p.stats) {
        const float vr = Traits<T>::round(v);
        const int db = r.b - b_first;
        if (db == 0) {
          ps0[j] += vr; pq0[j] = fmaf(vr, vr, pq0[j]);
        } else if (db == 1) {
          ps1[j] += vr; pq1[j] = fmaf(vr, vr, pq1[j]);
        } else {  // a 128-voxel tile spanning more than two samples (tiny volumes only)
          double* st = p.stats + ((long long)r.b * p.Cout + co) * 2;
          atomicAdd(st, (double)vr);
          atomicAdd(st + 1, (double)vr * vr);
        }
      }
    }
  }
  if (p.stats) {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      s_part[0][0][tm][tn * TN + j] = ps0[j]; s_part[0][1][tm][tn * TN + j] = ps1[j];
      s_part[1][0][tm][tn * TN + j] = pq0[j]; s_part[1][1][tm][tn * TN + j] = pq1[j];
    }
    __syncthreads();
    if (tid < 2 * BN) {
      const int db = tid / BN, c = tid % BN;
      const int co = n0 + c, b = b_first + db;
      float su = 0.f, sq = 0.f;
#pragma unroll
      for (int t = 0; t < 16; ++t) { su += s_part[0][db][t][c]; sq += s_part[1][db][t][c]; }
      if (co < p.Cout && b >= 0 && b < p.B && (su != 0.f || sq != 0.f)) {
        double* st = p.stats + ((long long)b * p.Cout + co) * 2;
        atomicAdd(st, (double)su);
        atomicAdd(st + 1, (double)sq);
      }
    }
  }
}

template <typename T, typename WT>
static int launch_conv_ffma(const mtb200_conv_params& p, cudaStream_t s) {
  const long long M = (long long)p.B * p.Do * p.Ho * p.Wo;
  if (M == 0) return MTB200_OK;
  if (p.Cout <= 32) {
    dim3 grid((unsigned)((M + FF_BM - 1) / FF_BM), (p.Cout + 31) / 32, p.ngroups);
    conv_taps_ffma_kernel<T, WT, 32><<<grid, FF_THREADS, 0, s>>>(p);
  } else {
    dim3 grid((unsigned)((M + FF_BM - 1) / FF_BM), (p.Cout + 63) / 64, p.ngroups);
    conv_taps_ffma_kernel<T, WT, 64><<<grid, FF_THREADS, 0, s>>>(p);
  }
  return check_launch("conv_taps_ffma");
}

int conv_taps_ffma(const mtb200_conv_params& p, cudaStream_t s) {
  MTB_REQUIRE(p.Cin % FF_KC == 0, "conv_taps(ffma): Cin=%d must be a multiple of 16", p.Cin);
  MTB_REQUIRE(p.in_ldc % 8 == 0 && p.in_coff % 8 == 0, "conv_taps(ffma): input channel stride/offset must be x8");
  if (p.dtype == MTB200_F32) {
    MTB_REQUIRE(p.wdtype == MTB200_F32, "conv_taps(ffma): f32 activations need f32 weights");
    return launch_conv_ffma<float, float>(p, s);
  } else if (p.dtype == MTB200_BF16) {
    if (p.wdtype == MTB200_BF16) return launch_conv_ffma<__nv_bfloat16, __nv_bfloat16>(p, s);
    if (p.wdtype == MTB200_F32) return launch_conv_ffma<__nv_bfloat16, float>(p, s);
  } else if (p.dtype == MTB200_F16) {
    if (p.wdtype == MTB200_F16) return launch_conv_ffma<__half, __half>(p, s);
    if (p.wdtype == MTB200_F32) return launch_conv_ffma<__half, float>(p, s);
  }
  set_error("conv_taps(ffma): unsupported dtype combination %d/%d", p.dtype, p.wdtype);
  return MTB200_ERR_INVALID;
}

// ---------------------------------------------------------------------------------------------------------------------
// weight gradient: dw[widx_t][co][ci] += sum_v dy[v][co] * f(x[v + off_t][ci])
// grid: x = split over voxels, y = (co tile, ci tile), z = tap
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WG_T = 64;   // co tile == ci tile
constexpr int WG_KC = 16;  // voxels per step

template <typename T>
__global__ void __launch_bounds__(FF_THREADS) wgrad_taps_ffma_kernel(const __grid_constant__ mtb200_wgrad_params p,
                                                                     long long chunk) {
  __shared__ __align__(16) float Ds[WG_KC][WG_T];
  __shared__ __align__(16) float Xs[WG_KC][WG_T];
  const int tid = threadIdx.x;
  const int t = blockIdx.z;
  int g = 0;
  while (t >= p.group_tap_begin[g + 1]) ++g;
  const int n_ci_tiles = (p.Cin + WG_T - 1) / WG_T;
  const int co0 = (blockIdx.y / n_ci_tiles) * WG_T, ci0 = (blockIdx.y % n_ci_tiles) * WG_T;
  const long long M = (long long)p.B * p.Do * p.Ho * p.Wo;
  const long long mbeg = (long long)blockIdx.x * chunk;
  const long long mend = min(M, mbeg + chunk);

  const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
  const T* __restrict__ dy = reinterpret_cast<const T*>(p.dy);

  const int lv = tid / 16, lc = (tid % 16) * 4;  // loader: voxel lv, 4 channels at lc
  const int tm = tid / 16, tn = tid % 16;        // compute: co rows tm*4..+4, ci cols tn*4..+4
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long mb = mbeg; mb < mend; mb += WG_KC) {
    const long long m = mb + lv;
    float dv[4] = {0.f, 0.f, 0.f, 0.f}, xv[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < mend) {
      const RowInfo r = decode_row(m, p.Do, p.Ho, p.Wo);
      const long long ovox = (((long long)r.b * p.Dof + (r.d * p.os[0] + p.group_ooff[g][0])) * p.Hof +
                              (r.h * p.os[1] + p.group_ooff[g][1])) * p.Wof + (r.w * p.os[2] + p.group_ooff[g][2]);
      if (co0 + lc < p.Cout) {
        const T* dp = dy + ovox * p.out_ldc + p.out_coff + co0 + lc;
#pragma unroll
        for (int j = 0; j < 4; ++j) dv[j] = Traits<T>::ld(dp + j);
      }
      const int id = r.d * p.is[0] + p.tap_off[t][0];
      const int ih = r.h * p.is[1] + p.tap_off[t][1];
      const int iw = r.w * p.is[2] + p.tap_off[t][2];
      if (ci0 + lc < p.Cin && (unsigned)id < (unsigned)p.Di && (unsigned)ih < (unsigned)p.Hi &&
          (unsigned)iw < (unsigned)p.Wi) {
        const long long vox = (((long long)r.b * p.Di + id) * p.Hi + ih) * p.Wi + iw;
        const T* xp = x + vox * p.in_ldc + p.in_coff + ci0 + lc;
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = Traits<T>::ld(xp + j);
        if (p.xform) {
          const float4* xf = reinterpret_cast<const float4*>(p.xform) + (long long)r.b * p.Cin + ci0 + lc;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 f = __ldg(xf + j);
            const float v = fmaf(xv[j], f.x, f.y);
            xv[j] = v > 0.f ? v : v * f.z;
          }
        }
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Ds[lv][lc]) = make_float4(dv[0], dv[1], dv[2], dv[3]);
    *reinterpret_cast<float4*>(&Xs[lv][lc]) = make_float4(xv[0], xv[1], xv[2], xv[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WG_KC; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&Ds[k][tm * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Xs[k][tn * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  float* dw = p.dw + (long long)p.tap_widx[t] * p.Cout * p.Cin;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + tm * 4 + i;
    if (co >= p.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tn * 4 + j;
      if (ci < p.Cin && acc[i][j] != 0.f) atomicAdd(dw + (long long)co * p.Cin + ci, acc[i][j]);
    }
  }
}

int wgrad_taps_ffma(const mtb200_wgrad_params& p, cudaStream_t s) {
  MTB_REQUIRE(p.Cin % 4 == 0 && p.Cout % 4 == 0, "wgrad_taps(ffma): channel counts must be multiples of 4");
  const long long M = (long long)p.B * p.Do * p.Ho * p.Wo;
  if (M == 0) return MTB200_OK;
  const int tiles = ((p.Cout + WG_T - 1) / WG_T) * ((p.Cin + WG_T - 1) / WG_T);
  const long long target_ctas = 4LL * num_sms();
  long long split = max(1LL, target_ctas / ((long long)tiles * p.ntaps));
  long long chunk = (M + split - 1) / split;
  chunk = max((long long)256, ((chunk + WG_KC - 1) / WG_KC) * WG_KC);
  split = (M + chunk - 1) / chunk;
  dim3 grid((unsigned)split, tiles, p.ntaps);
  MTB_DISPATCH_DTYPE(p.dtype, T, (wgrad_taps_ffma_kernel<T><<<grid, FF_THREADS, 0, s>>>(p, chunk)));
  return check_launch("wgrad_taps_ffma");
}

// ---------------------------------------------------------------------------------------------------------------------
// column sums (bias gradients)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ m, long long rows, int ldc, int coff, int C, float* out,
                              long long rows_per_block) {
  // block = 256 threads = 8 warps; lane -> channel (c0 + lane), warps stride over rows
  const int c = blockIdx.y * 32 + (threadIdx.x & 31);
  const int wid = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float acc = 0.f;
  if (c < C)
    for (long long r = r0 + wid; r < r1; r += 8) acc += Traits<T>::ld(m + r * ldc + coff + c);
  __shared__ float sh[8][32];
  sh[wid][threadIdx.x & 31] = acc;
  __syncthreads();
  if (wid == 0 && c < C) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += sh[i][threadIdx.x];
    atomicAdd(out + c, v);
  }
}

int colsum(const void* m, int dtype, long long rows, int ldc, int coff, int C, float* out, cudaStream_t s) {
  if (rows == 0 || C == 0) return MTB200_OK;
  long long nblk = min((long long)num_sms() * 4, (rows + 255) / 256);
  long long rpb = (rows + nblk - 1) / nblk;
  dim3 grid((unsigned)nblk, (C + 31) / 32);
  MTB_DISPATCH_DTYPE(dtype, T, (colsum_kernel<T><<<grid, 256, 0, s>>>(reinterpret_cast<const T*>(m), rows, ldc, coff,
                                                                      C, out, rpb)));
  return check_launch("colsum");
}

}  // namespace mtb
