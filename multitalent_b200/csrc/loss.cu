// MultiTalent multi-head loss: sigmoid + BCE-with-logits (mean over voxels per (sample, region)) - soft Dice on the
// sigmoid pooled over ranks per (local batch index, channel).  Two streaming passes over the logits (HBM-bound):
// pass 1 = statistics, pass 2 = d(loss)/d(logits).  Replaces the python double loop at
// MultiTalent_Trainer_DDP.py:567-606 and its autograd graph (~1e3 tiny kernel launches per step in the reference).
#include "common.cuh"

namespace mtb {

constexpr int LT = 256;
constexpr int MAX_LABELS = 64;

// e = exp(-|z|), sigmoid(z) and softplus(-|z|) = log(1 + e) on the SFU (ex2 / rcp / lg2 approximations, ~1e-6 relative:
// far inside the fp32 parity tolerance of the loss).  The precise expf / division / log1pf sequences made both passes
// instruction-bound (lanes of one warp own different channel groups, so every channel's code is issued for the whole warp).
__device__ __forceinline__ void sigmoid_terms(float z, float& e, float& sig) {
  e = __expf(-fabsf(z));
  const float r = __frcp_rn(1.f + e);
  sig = z >= 0.f ? r : e * r;
}

// HARD: additionally count, per supervised (b, channel), the thresholded prediction (sigmoid(z) > 0.5 <=> z > 0) against
// the same region target: hard[b][c] = {sum pred * y, sum pred} -- with sum y from the soft statistics that is the
// tp / fp / fn of run_online_evaluation (MultiTalent_Trainer_DDP.py:372-397) out of the pass the loss makes anyway.
template <typename T, bool HARD>
__global__ void __launch_bounds__(LT) mt_loss_stats_kernel(const T* __restrict__ logits, int ldc, int C,
                                                           const float* __restrict__ target, long long nvox,
                                                           const uint64_t* __restrict__ valid_mask,
                                                           const uint64_t* __restrict__ pos_mask, int n_labels,
                                                           double* __restrict__ stats, double* __restrict__ hard) {
  constexpr int NQ = HARD ? 6 : 4;
  __shared__ uint64_t s_pos[MAX_LABELS];
  __shared__ float sh[LT][8];
  __shared__ int s_act[8], s_nact;
  const int b = blockIdx.y;
  const int G = C / 8;
  const uint64_t valid = valid_mask[b];
  // only channel groups with a supervised channel are read: the threads of the block are spread over THOSE groups
  // (a sample of a 1..13-region dataset touches 1..3 of the 6 groups; idle threads would only thin out the loads in flight)
  if (threadIdx.x == 0) {
    int n = 0;
    for (int g = 0; g < G; ++g)
      if ((valid >> (g * 8)) & 0xffull) s_act[n++] = g;
    s_nact = n;
  }
  if (threadIdx.x < MAX_LABELS) s_pos[threadIdx.x] = threadIdx.x < n_labels ? pos_mask[threadIdx.x] : 0ull;
  __syncthreads();
  const int nact = s_nact;
  if (nact == 0) return;  // nothing supervised in this sample (uniform across the block)
  const int vstride = LT / nact;
  const int ai = threadIdx.x % nact, vlane = threadIdx.x / nact;
  const int cg = s_act[ai];
  const bool active = vlane < vstride;
  const unsigned vbits = (unsigned)((valid >> (cg * 8)) & 0xffull);
  float part[NQ][8];
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int j = 0; j < 8; ++j) part[q][j] = 0.f;

  if (active) {
    const T* base = logits + (long long)b * nvox * ldc + cg * 8;
    const float* tb = target + (long long)b * nvox;
    constexpr int U = 4;  // voxels in flight per thread
    const long long per = (nvox + gridDim.x - 1) / gridDim.x;
    const long long v0 = (long long)blockIdx.x * per, v1 = min(nvox, v0 + per);
    for (long long v = v0 + vlane; v < v1; v += (long long)U * vstride) {
      float z[U][8];
      int lab[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long vv = v + (long long)u * vstride;
        if (vv < v1) {
          load8<T>(base + vv * ldc, z[u]);
          lab[u] = (int)tb[vv];
        } else {
          lab[u] = -1;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (lab[u] < 0) continue;   // past the end (labels are >= 0)
        const uint64_t pm = ((unsigned)lab[u] < (unsigned)MAX_LABELS) ? s_pos[lab[u]] : 0ull;
        const unsigned ybits = (unsigned)((pm >> (cg * 8)) & 0xffull);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (vbits & (1u << j)) {
            const float zz = z[u][j];
            const float y = (ybits >> j) & 1u ? 1.f : 0.f;
            float e, sig;
            sigmoid_terms(zz, e, sig);
            // log1p(e): e < 2^-12 -> e - e^2/2 (log(1+e) would lose the low bits of e in the rounding of 1 + e)
            const float l1p = e < 2.44140625e-4f ? e * (1.f - 0.5f * e) : __logf(1.f + e);
            const float bce = fmaxf(zz, 0.f) - zz * y + l1p;
            part[0][j] += bce;
            part[1][j] = fmaf(sig, y, part[1][j]);
            part[2][j] += sig;
            part[3][j] += y;
            if (HARD) {
              const float pred = zz > 0.f ? 1.f : 0.f;
              part[NQ - 2][j] += pred * y;
              part[NQ - 1][j] += pred;
            }
          }
        }
      }
    }
  }
  // block reduce over the voxel lanes, one double atomic per (channel, stat)
  double* dst = stats + (long long)b * C * 4;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[threadIdx.x][j] = active ? part[q][j] : 0.f;
    __syncthreads();
    for (int idx = threadIdx.x; idx < nact * 8; idx += LT) {
      const int a = idx / 8, j = idx % 8;
      float sum = 0.f;
      for (int vl = 0; vl < vstride; ++vl) sum += sh[vl * nact + a][j];
      if (sum != 0.f) {
        if (q < 4) atomicAdd(dst + (long long)(s_act[a] * 8 + j) * 4 + q, (double)sum);
        else atomicAdd(hard + ((long long)b * C + s_act[a] * 8 + j) * 2 + (q - 4), (double)sum);
      }
    }
  }
}

static dim3 loss_grid(long long nvox, int B, int C) {
  const int vstride = LT / (C / 8);
  long long want = (8LL * num_sms() + B - 1) / B;
  long long maxb = (nvox + (long long)vstride * 8 - 1) / ((long long)vstride * 8);
  return dim3((unsigned)max(1LL, min(want, maxb)), (unsigned)B);
}

int mt_loss_stats(const void* logits, int dtype, int ldc, int C, const float* target, int B, long long nvox,
                  const uint64_t* valid_mask, const uint64_t* pos_mask, int n_labels, double* stats, double* hard,
                  cudaStream_t s) {
  MTB_REQUIRE(C % 8 == 0 && C <= 64 && ldc % 8 == 0 && C <= ldc, "mt_loss_stats: C=%d (padded, <=64) ldc=%d", C, ldc);
  MTB_REQUIRE(n_labels <= MAX_LABELS, "mt_loss_stats: n_labels=%d > %d", n_labels, MAX_LABELS);
  dim3 grid = loss_grid(nvox, B, C);
  if (hard) {
    MTB_DISPATCH_DTYPE(dtype, T, (mt_loss_stats_kernel<T, true><<<grid, LT, 0, s>>>(
        reinterpret_cast<const T*>(logits), ldc, C, target, nvox, valid_mask, pos_mask, n_labels, stats, hard)));
  } else {
    MTB_DISPATCH_DTYPE(dtype, T, (mt_loss_stats_kernel<T, false><<<grid, LT, 0, s>>>(
        reinterpret_cast<const T*>(logits), ldc, C, target, nvox, valid_mask, pos_mask, n_labels, stats, nullptr)));
  }
  return check_launch("mt_loss_stats");
}

// ---- finalize (tiny): loss scalars and per-(b,j) gradient coefficients --------------------------------------------
__global__ void mt_loss_finalize_kernel(const double* __restrict__ stats, const double* __restrict__ pooled,
                                        const uint64_t* __restrict__ valid_mask, int B, int C, double inv_nvox,
                                        float weight, float world, float* __restrict__ losses,
                                        float4* __restrict__ coef) {
  pdl_wait();
  __shared__ double s_ce[256], s_dc[256];
  double ce = 0.0, dc = 0.0;
  for (int i = threadIdx.x; i < B * C; i += blockDim.x) {
    const int b = i / C, j = i % C;
    const bool valid = (valid_mask[b] >> j) & 1ull;
    const double TP = pooled ? pooled[2 * i] : stats[4 * i + 1];
    const double D = pooled ? pooled[2 * i + 1] : stats[4 * i + 2] + stats[4 * i + 3];
    const double Dc = D < 1e-7 ? 1e-7 : D;
    dc += 2.0 * TP / Dc;
    float c1 = 0.f, c2 = 0.f;
    if (valid) {
      ce += stats[4 * i] * inv_nvox;
      c1 = (float)((double)weight * world * 2.0 / Dc);
      c2 = D < 1e-7 ? 0.f : (float)((double)weight * world * 2.0 * TP / (D * D));
    }
    coef[i] = make_float4(valid ? (float)(weight * inv_nvox) : 0.f, c1, c2, valid ? 1.f : 0.f);
  }
  s_ce[threadIdx.x] = ce;
  s_dc[threadIdx.x] = dc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, d = 0.0;
    for (int i = 0; i < blockDim.x; ++i) { a += s_ce[i]; d += s_dc[i]; }
    losses[0] += weight * (float)(a - d);
    losses[1] += weight * (float)a;
    losses[2] += weight * (float)d;
  }
}

int mt_loss_finalize(const double* stats, const double* pooled, const uint64_t* valid_mask, int B, int C, long long nvox,
                     float weight, float world_size, float* losses, float* coef, cudaStream_t s) {
  MTB_REQUIRE(C <= 64, "mt_loss_finalize: C=%d", C);
  launch_pdl(mt_loss_finalize_kernel, dim3(1), dim3(256), (size_t)(0), s, stats, pooled, valid_mask, B, C, 1.0 / (double)nvox, weight, world_size,
                                            losses, reinterpret_cast<float4*>(coef));
  return check_launch("mt_loss_finalize");
}

// ---- pass 2: dlogits ----------------------------------------------------------------------------------------------
// Measured (profiles/r1j_ncu_full_loss.txt): 3.0 GB of DRAM traffic in 0.83 ms at full resolution.  More loads in flight
// (2 / 4 voxels per thread), coefficients in shared memory, SFU math and a chunked grid-stride order were all tried and
// none moved it (0.97 -> 1.04 / 1.13 / 1.05 / 1.40 ms per step): left as the simple one-voxel loop.
template <typename T>
__global__ void __launch_bounds__(LT) mt_loss_bwd_kernel(const T* __restrict__ logits, int ldc, int C,
                                                         const float* __restrict__ target, long long nvox,
                                                         const uint64_t* __restrict__ pos_mask, int n_labels,
                                                         const float4* __restrict__ coef,
                                                         const float* __restrict__ gscale, T* __restrict__ dlogits,
                                                         int d_ldc) {
  __shared__ uint64_t s_pos[MAX_LABELS];
  const int b = blockIdx.y;
  const int G = C / 8;
  const int vstride = LT / G;
  const int cg = threadIdx.x % G, vlane = threadIdx.x / G;
  if (threadIdx.x < MAX_LABELS) s_pos[threadIdx.x] = threadIdx.x < n_labels ? pos_mask[threadIdx.x] : 0ull;
  __syncthreads();
  if (vlane >= vstride) return;
  const float gs = gscale ? *gscale : 1.f;
  float4 cf[8];
  unsigned vbits = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    cf[j] = coef[(long long)b * C + cg * 8 + j];
    if (cf[j].w != 0.f) vbits |= 1u << j;
    cf[j].x *= gs; cf[j].y *= gs; cf[j].z *= gs;
  }
  const long long per = (nvox + gridDim.x - 1) / gridDim.x;
  const long long v0 = (long long)blockIdx.x * per, v1 = min(nvox, v0 + per);
  const T* base = logits + (long long)b * nvox * ldc + cg * 8;
  T* obase = dlogits + (long long)b * nvox * d_ldc + cg * 8;
  const float* tb = target + (long long)b * nvox;
  for (long long v = v0 + vlane; v < v1; v += vstride) {
    float d[8];
    if (vbits) {
      float z[8];
      load8<T>(base + v * ldc, z);
      const int lab = (int)tb[v];
      const uint64_t pm = ((unsigned)lab < (unsigned)MAX_LABELS) ? s_pos[lab] : 0ull;
      const unsigned ybits = (unsigned)((pm >> (cg * 8)) & 0xffull);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (vbits & (1u << j)) {
          const float y = (ybits >> j) & 1u ? 1.f : 0.f;
          float e, sig;
          sigmoid_terms(z[j], e, sig);
          d[j] = cf[j].x * (sig - y) - sig * (1.f - sig) * (y * cf[j].y - cf[j].z);
        } else {
          d[j] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = 0.f;
    }
    store8<T>(obase + v * d_ldc, d);
  }
}

int mt_loss_bwd(const void* logits, int dtype, int ldc, int C, const float* target, int B, long long nvox,
                const uint64_t* pos_mask, int n_labels, const float* coef, const float* gscale, void* dlogits, int d_ldc,
                cudaStream_t s) {
  MTB_REQUIRE(C % 8 == 0 && C <= 64 && ldc % 8 == 0 && d_ldc % 8 == 0 && C <= ldc && C <= d_ldc,
              "mt_loss_bwd: C=%d ldc=%d d_ldc=%d", C, ldc, d_ldc);
  MTB_REQUIRE(n_labels <= MAX_LABELS, "mt_loss_bwd: n_labels=%d > %d", n_labels, MAX_LABELS);
  dim3 grid = loss_grid(nvox, B, C);
  MTB_DISPATCH_DTYPE(dtype, T, (mt_loss_bwd_kernel<T><<<grid, LT, 0, s>>>(
      reinterpret_cast<const T*>(logits), ldc, C, target, nvox, pos_mask, n_labels,
      reinterpret_cast<const float4*>(coef), gscale, reinterpret_cast<T*>(dlogits), d_ldc)));
  return check_launch("mt_loss_bwd");
}

}  // namespace mtb
