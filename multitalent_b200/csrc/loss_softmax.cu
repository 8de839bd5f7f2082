// Generic nnU-Net loss: softmax + cross entropy + soft Dice (DC_and_CE_loss, training/loss_functions/dice_loss.py:488-545;
// SoftDiceLoss :155-195; get_tp_fp_fn_tn :100-152; RobustCrossEntropyLoss crossentropy.py:4-11).  Two streaming passes
// over the NDHWC logits (HBM-bound): pass 1 = per-(sample, class) statistics {tp = sum p*onehot, Sp = sum p,
// N = sum onehot} and the per-sample cross-entropy sum; pass 2 = d(loss)/d(logits) from per-class coefficients.
//
// Thread mapping: a voxel is owned by GP = next_pow2(C/8) neighbouring lanes, 8 channels each; the softmax max / sum are
// width-GP shuffle reductions, so no thread ever holds more than 8 channels.
#include "common.cuh"

namespace mtb {

constexpr int ST = 256;

template <typename T, int GP>
__global__ void __launch_bounds__(ST) dcce_stats_kernel(const T* __restrict__ logits, int ldc, int C /* real classes */,
                                                        const float* __restrict__ target, long long nvox,
                                                        double* __restrict__ stats /* [B][Cp][3] */,
                                                        double* __restrict__ ce_sum /* [B] */, int Cp) {
  __shared__ float sh[ST][8];
  const int b = blockIdx.y;
  const int cg = threadIdx.x % GP, vlane = threadIdx.x / GP;
  constexpr int VS = ST / GP;
  const bool has_ch = cg * 8 < Cp;
  const T* base = logits + (long long)b * nvox * ldc + cg * 8;
  const float* tb = target + (long long)b * nvox;
  float tp[8], sp[8], cnt[8], ce = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) tp[j] = sp[j] = cnt[j] = 0.f;
  // block-uniform trip count (the shuffles below need every lane of the warp); tail lanes re-read the last voxel
  for (long long v0 = (long long)blockIdx.x * VS; v0 < nvox; v0 += (long long)gridDim.x * VS) {
    const bool live = v0 + vlane < nvox;
    const long long v = live ? v0 + vlane : nvox - 1;
    float z[8];
    if (has_ch) load8<T>(base + v * ldc, z);
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (!has_ch || cg * 8 + j >= C) z[j] = -INFINITY;
      m = fmaxf(m, z[j]);
    }
#pragma unroll
    for (int o = GP / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o, GP));
    float e[8], s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { e[j] = __expf(z[j] - m); s += e[j]; }
#pragma unroll
    for (int o = GP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, GP);
    const float inv = 1.f / s;
    const int y = (int)tb[v];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p = live ? e[j] * inv : 0.f;
      sp[j] += p;
      if (live && cg * 8 + j == y) {
        tp[j] += p;
        cnt[j] += 1.f;
        ce += logf(s) - (z[j] - m);  // -log softmax_y
      }
    }
  }
  // block reduction over the voxel lanes
  double* dst = stats + (long long)b * Cp * 3;
  for (int q = 0; q < 4; ++q) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[threadIdx.x][j] = q == 0 ? tp[j] : (q == 1 ? sp[j] : (q == 2 ? cnt[j] : (j == 0 ? ce : 0.f)));
    __syncthreads();
    for (int idx = threadIdx.x; idx < GP * 8; idx += ST) {
      const int g = idx / 8, j = idx % 8;
      float a = 0.f;
      for (int vl = 0; vl < VS; ++vl) a += sh[vl * GP + g][j];
      if (q < 3) {
        if (g * 8 + j < Cp && a != 0.f) atomicAdd(dst + (long long)(g * 8 + j) * 3 + q, (double)a);
      } else if (j == 0 && a != 0.f) {
        atomicAdd(ce_sum + b, (double)a);
      }
    }
  }
}

// d(loss)/dz_k = gscale * [ ce_w * (p_k - [k == y]) + p_k * (G_k - sum_c G_c p_c) ],  G_c = coef[b][c][0] * [c == y] + coef[b][c][1]
template <typename T, int GP>
__global__ void __launch_bounds__(ST) dcce_bwd_kernel(const T* __restrict__ logits, int ldc, int C,
                                                      const float* __restrict__ target, long long nvox,
                                                      const float* __restrict__ coef /* [B][Cp][2] */, float ce_w,
                                                      const float* __restrict__ gscale, T* __restrict__ dz, int dz_ldc,
                                                      int Cp) {
  const int b = blockIdx.y;
  const int cg = threadIdx.x % GP, vlane = threadIdx.x / GP;
  constexpr int VS = ST / GP;
  const bool has_ch = cg * 8 < Cp;
  const T* base = logits + (long long)b * nvox * ldc + cg * 8;
  T* obase = dz + (long long)b * nvox * dz_ldc + cg * 8;
  const float* tb = target + (long long)b * nvox;
  float ca[8], cb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    ca[j] = (has_ch && c < C) ? coef[((long long)b * Cp + c) * 2] : 0.f;
    cb[j] = (has_ch && c < C) ? coef[((long long)b * Cp + c) * 2 + 1] : 0.f;
  }
  const float gs = gscale[0];
  // block-uniform trip count (the shuffles below need every lane of the warp); tail lanes re-read the last voxel
  for (long long v0 = (long long)blockIdx.x * VS; v0 < nvox; v0 += (long long)gridDim.x * VS) {
    const bool live = v0 + vlane < nvox;
    const long long v = live ? v0 + vlane : nvox - 1;
    float z[8];
    if (has_ch) load8<T>(base + v * ldc, z);
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (!has_ch || cg * 8 + j >= C) z[j] = -INFINITY;
      m = fmaxf(m, z[j]);
    }
#pragma unroll
    for (int o = GP / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o, GP));
    float p[8], s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { p[j] = __expf(z[j] - m); s += p[j]; }
#pragma unroll
    for (int o = GP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, GP);
    const float inv = 1.f / s;
    const int y = (int)tb[v];
    float G[8], dot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      p[j] *= inv;
      G[j] = cb[j] + ((cg * 8 + j == y) ? ca[j] : 0.f);
      dot = fmaf(G[j], p[j], dot);
    }
#pragma unroll
    for (int o = GP / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o, GP);
    if (has_ch && live) {
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float oh = (cg * 8 + j == y) ? 1.f : 0.f;
        o8[j] = (cg * 8 + j < C) ? gs * (ce_w * (p[j] - oh) + p[j] * (G[j] - dot)) : 0.f;
      }
      store8<T>(obase + v * dz_ldc, o8);
    }
  }
}

static dim3 dcce_grid(long long nvox, int B, int gp) {
  const int vs = ST / gp;
  long long want = (8LL * num_sms() + B - 1) / B;
  long long maxb = (nvox + vs - 1) / vs;
  return dim3((unsigned)max(1LL, min(want, maxb)), (unsigned)B);
}

#define DCCE_DISPATCH(KERNEL, ...)                                                            \
  MTB_DISPATCH_DTYPE(dtype, T, {                                                              \
    if (gp == 1) KERNEL<T, 1><<<grid, ST, 0, s>>>(__VA_ARGS__);                               \
    else if (gp == 2) KERNEL<T, 2><<<grid, ST, 0, s>>>(__VA_ARGS__);                          \
    else if (gp == 4) KERNEL<T, 4><<<grid, ST, 0, s>>>(__VA_ARGS__);                          \
    else KERNEL<T, 8><<<grid, ST, 0, s>>>(__VA_ARGS__);                                       \
  })

int dcce_stats(const void* logits, int dtype, int ldc, int C, int Cp, const float* target, int B, long long nvox,
               double* stats, double* ce_sum, cudaStream_t s) {
  MTB_REQUIRE(Cp % 8 == 0 && Cp <= 64 && C <= Cp && ldc % 8 == 0 && ldc >= Cp, "dcce_stats: C=%d Cp=%d ldc=%d", C, Cp, ldc);
  if (nvox == 0 || B == 0) return MTB200_OK;
  int gp = 1;
  while (gp * 8 < Cp) gp *= 2;
  dim3 grid = dcce_grid(nvox, B, gp);
  DCCE_DISPATCH(dcce_stats_kernel, reinterpret_cast<const T*>(logits), ldc, C, target, nvox, stats, ce_sum, Cp);
  return check_launch("dcce_stats");
}

int dcce_bwd(const void* logits, int dtype, int ldc, int C, int Cp, const float* target, int B, long long nvox,
             const float* coef, float ce_w, const float* gscale, void* dz, int dz_ldc, cudaStream_t s) {
  MTB_REQUIRE(Cp % 8 == 0 && Cp <= 64 && C <= Cp && ldc % 8 == 0 && ldc >= Cp && dz_ldc % 8 == 0 && dz_ldc >= Cp,
              "dcce_bwd: C=%d Cp=%d ldc=%d dz_ldc=%d", C, Cp, ldc, dz_ldc);
  if (nvox == 0 || B == 0) return MTB200_OK;
  int gp = 1;
  while (gp * 8 < Cp) gp *= 2;
  dim3 grid = dcce_grid(nvox, B, gp);
  DCCE_DISPATCH(dcce_bwd_kernel, reinterpret_cast<const T*>(logits), ldc, C, target, nvox, coef, ce_w, gscale,
                reinterpret_cast<T*>(dz), dz_ldc, Cp);
  return check_launch("dcce_bwd");
}

}  // namespace mtb
