// InstanceNorm3d(affine) + LeakyReLU: statistics finalisation, materialisation, and the two-pass backward.
// All kernels are HBM-bound streaming passes over NDHWC slices: one thread owns a group of 8 channels (one 16/32-byte
// vector) and strides over voxels, so the per-(b,c) coefficients live in registers.
#include <stdlib.h>

#include "common.cuh"

namespace mtb {

constexpr int NT = 256;

// norm_tma.cu
int in_bwd_apply_tma(const void* dact, int d_ldc, int d_coff, const void* y, int y_ldc, int y_coff, void* dy, int dy_ldc,
                     int dy_coff, int dtype, int B, long long nvox, int C, const float* xform, const float* meanrstd,
                     const float* gamma, const double* red, float* dgamma, float* dbeta, cudaStream_t s);
int norm_act_tma(const void* y, int in_ldc, int in_coff, void* out, int out_ldc, int out_coff, int dtype, int B,
                 long long nvox, int C, const float* xform, cudaStream_t s);
int in_bwd_reduce_tma(const void* dact, int d_ldc, int d_coff, const void* y, int y_ldc, int y_coff, int dtype, int B,
                      long long nvox, int C, const float* xform, const float* meanrstd, double* red, cudaStream_t s);

struct Span {  // work split of one sample's voxels over gridDim.x blocks
  long long v0, v1;
  int cg, vlane, vstride;
  bool active;
};

__device__ __forceinline__ Span make_span(long long nvox, int C) {
  const int G = C / 8;
  Span s;
  s.vstride = NT / G;
  s.cg = threadIdx.x % G;
  s.vlane = threadIdx.x / G;
  s.active = s.vlane < s.vstride;
  const long long per = (nvox + gridDim.x - 1) / gridDim.x;
  s.v0 = (long long)blockIdx.x * per;
  s.v1 = min(nvox, s.v0 + per);
  return s;
}

// tuning knobs (environment overrides are for tools/norm_bench.py only; defaults = what measured best on the B200)
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
static int norm_nu() { static int v = env_int("MTB200_NORM_NU", 4); return v; }
static int norm_waves() { static int v = env_int("MTB200_NORM_WAVES", 8); return v; }

static dim3 span_grid(long long nvox, int B, int C) {
  const int G = C / 8;
  const int vstride = NT / G;
  long long want = ((long long)norm_waves() * num_sms() + B - 1) / B;
  long long maxb = (nvox + (long long)vstride * 4 - 1) / ((long long)vstride * 4);  // >= 4 iterations per block
  long long nb = max(1LL, min(want, maxb));
  return dim3((unsigned)nb, (unsigned)B);
}

// ---- finalize: stats -> {scale, shift, slope} ----------------------------------------------------------------------
__global__ void in_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int B, int C, double inv_n, float eps, float slope,
                                   float4* __restrict__ xform, float2* __restrict__ meanrstd) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C;
  const double mean = stats[2 * i] * inv_n;
  double var = stats[2 * i + 1] * inv_n - mean * mean;  // biased variance, as torch.instance_norm
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * rstd;
  xform[i] = make_float4(sc, beta[c] - (float)mean * sc, slope, 0.f);
  meanrstd[i] = make_float2((float)mean, rstd);
}

int in_finalize(const double* stats, const float* gamma, const float* beta, int B, int C, long long nvox, float eps,
                float slope, float* xform, float* meanrstd, cudaStream_t s) {
  const int n = B * C;
  launch_pdl(in_finalize_kernel, dim3((n + 127) / 128), dim3(128), (size_t)(0), s, stats, gamma, beta, B, C, 1.0 / (double)nvox, eps, slope,
                                                     reinterpret_cast<float4*>(xform),
                                                     reinterpret_cast<float2*>(meanrstd));
  return check_launch("in_finalize");
}

// ---- block reduction of 8-channel partials over the threads sharing a channel group -------------------------------
template <int NV>
__device__ __forceinline__ void reduce_store(float (&part)[NV][8], const Span& sp, int C, double* dst /* [C][NV] */) {
  __shared__ float sh[NT][8];
  const int G = C / 8;
  for (int q = 0; q < NV; ++q) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[threadIdx.x][j] = sp.active ? part[q][j] : 0.f;
    __syncthreads();
    // thread (cg, j) for vlane == 0 sums over vlanes
    for (int idx = threadIdx.x; idx < G * 8; idx += NT) {
      const int cg = idx / 8, j = idx % 8;
      float v = 0.f;
      for (int vl = 0; vl < sp.vstride; ++vl) v += sh[vl * G + cg][j];
      if (v != 0.f) atomicAdd(dst + (long long)(cg * 8 + j) * NV + q, (double)v);
    }
  }
}

// ---- plain statistics pass ----------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT) in_stats_kernel(const T* __restrict__ y, long long nvox, int ldc, int coff, int C,
                                                      double* __restrict__ stats) {
  const Span sp = make_span(nvox, C);
  const int b = blockIdx.y;
  float part[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) part[0][j] = part[1][j] = 0.f;
  if (sp.active) {
    const T* base = y + (long long)b * nvox * ldc + coff + sp.cg * 8;
    for (long long v = sp.v0 + sp.vlane; v < sp.v1; v += sp.vstride) {
      float x[8];
      load8<T>(base + v * ldc, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) { part[0][j] += x[j]; part[1][j] = fmaf(x[j], x[j], part[1][j]); }
    }
  }
  reduce_store<2>(part, sp, C, stats + (long long)b * C * 2);
}

int in_stats(const void* y, int dtype, int B, long long nvox, int ldc, int coff, int C, double* stats, cudaStream_t s) {
  MTB_REQUIRE(C % 8 == 0 && C / 8 <= NT && ldc % 8 == 0 && coff % 8 == 0, "in_stats: C=%d ldc=%d coff=%d", C, ldc, coff);
  dim3 grid = span_grid(nvox, B, C);
  MTB_DISPATCH_DTYPE(dtype, T, (in_stats_kernel<T><<<grid, NT, 0, s>>>(reinterpret_cast<const T*>(y), nvox, ldc, coff,
                                                                       C, stats)));
  return check_launch("in_stats");
}

// ---- materialise act = f(y) (+ residual) ---------------------------------------------------------------------------
template <typename T, int NU>
__global__ void __launch_bounds__(NT) norm_act_kernel(const T* __restrict__ y, int in_ldc, int in_coff, T* __restrict__ out,
                                                      int out_ldc, int out_coff, long long nvox, int C,
                                                      const float4* __restrict__ xform, const T* __restrict__ res,
                                                      int res_ldc, int res_coff, const float4* __restrict__ res_xform,
                                                      float slope2) {
  const Span sp = make_span(nvox, C);
  if (!sp.active) return;
  const int b = blockIdx.y;
  float4 f[8], rf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    f[j] = xform ? xform[(long long)b * C + sp.cg * 8 + j] : make_float4(1.f, 0.f, 1.f, 0.f);
    rf[j] = res_xform ? res_xform[(long long)b * C + sp.cg * 8 + j] : make_float4(1.f, 0.f, 1.f, 0.f);
  }
  const T* ybase = y + (long long)b * nvox * in_ldc + in_coff + sp.cg * 8;
  T* obase = out + (long long)b * nvox * out_ldc + out_coff + sp.cg * 8;
  const T* rbase = res ? res + (long long)b * nvox * res_ldc + res_coff + sp.cg * 8 : nullptr;
  if (rbase) {
    for (long long v = sp.v0 + sp.vlane; v < sp.v1; v += sp.vstride) {
      float x[8], r[8];
      load8<T>(ybase + v * in_ldc, x);
      load8<T>(rbase + v * res_ldc, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = fmaf(x[j], f[j].x, f[j].y);
        t = t > 0.f ? t : t * f[j].z;
        float u = fmaf(r[j], rf[j].x, rf[j].y);
        u = u > 0.f ? u : u * rf[j].z;
        u += t;
        x[j] = u > 0.f ? u : u * slope2;
      }
      store8<T>(obase + v * out_ldc, x);
    }
    return;
  }
  // streaming pass: NU voxels per thread per iteration, every load issued before the arithmetic (memory-level
  // parallelism: ~6.5 MB must be in flight chip-wide to saturate HBM3e)
  const long long chunk = (long long)NU * sp.vstride;  // grid-stride over chunks: the active blocks form a compact
                                                       // moving front in DRAM instead of 1000+ scattered streams
  for (long long base = (long long)blockIdx.x * chunk; base < nvox; base += (long long)gridDim.x * chunk) {
    float x[NU][8];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const long long vv = base + sp.vlane + (long long)u * sp.vstride;
      if (vv < nvox) load8<T>(ybase + vv * in_ldc, x[u]);
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const long long vv = base + sp.vlane + (long long)u * sp.vstride;
      if (vv < nvox) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(x[u][j], f[j].x, f[j].y);
          x[u][j] = t > 0.f ? t : t * f[j].z;
        }
        store8<T>(obase + vv * out_ldc, x[u]);
      }
    }
  }
}

int norm_act(const void* y, int in_ldc, int in_coff, void* out, int out_ldc, int out_coff, int dtype, int B,
             long long nvox, int C, const float* xform, const void* res, int res_ldc, int res_coff,
             const float* res_xform, float slope2, cudaStream_t s) {
  MTB_REQUIRE(C % 8 == 0 && C / 8 <= NT && in_ldc % 8 == 0 && in_coff % 8 == 0 && out_ldc % 8 == 0 && out_coff % 8 == 0,
              "norm_act: channel counts/strides must be multiples of 8 (C=%d)", C);
  if (res) MTB_REQUIRE(res_ldc % 8 == 0 && res_coff % 8 == 0, "norm_act: residual stride/offset must be x8");
  if (!res && dtype != MTB200_F32) {
    const int r = norm_act_tma(y, in_ldc, in_coff, out, out_ldc, out_coff, dtype, B, nvox, C, xform, s);
    if (r != MTB200_ERR_UNSUPPORTED) return r;
  }
  dim3 grid = span_grid(nvox, B, C);
#define MTB_NORM_ACT(NU_)                                                                                        \
  MTB_DISPATCH_DTYPE(dtype, T, (norm_act_kernel<T, NU_><<<grid, NT, 0, s>>>(                                     \
      reinterpret_cast<const T*>(y), in_ldc, in_coff, reinterpret_cast<T*>(out), out_ldc, out_coff, nvox, C,     \
      reinterpret_cast<const float4*>(xform), reinterpret_cast<const T*>(res), res_ldc, res_coff,                \
      reinterpret_cast<const float4*>(res_xform), slope2)))
  switch (norm_nu()) { case 1: MTB_NORM_ACT(1); break; case 2: MTB_NORM_ACT(2); break; default: MTB_NORM_ACT(4); }
#undef MTB_NORM_ACT
  return check_launch("norm_act");
}

// ---- backward pass 1: red[b][c] = {sum dv, sum dv*xhat}, dv = dact * lrelu'(v) --------------------------------------
template <typename T, int NU>
__global__ void __launch_bounds__(NT) in_bwd_reduce_kernel(const T* __restrict__ dact, int d_ldc, int d_coff,
                                                           const T* __restrict__ y, int y_ldc, int y_coff, long long nvox,
                                                           int C, const float4* __restrict__ xform,
                                                           const float2* __restrict__ meanrstd, double* __restrict__ red) {
  const Span sp = make_span(nvox, C);
  const int b = blockIdx.y;
  float part[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) part[0][j] = part[1][j] = 0.f;
  if (sp.active) {
    float4 f[8];
    float2 mr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = xform[(long long)b * C + sp.cg * 8 + j];
      mr[j] = meanrstd[(long long)b * C + sp.cg * 8 + j];
    }
    const T* dbase = dact + (long long)b * nvox * d_ldc + d_coff + sp.cg * 8;
    const T* ybase = y + (long long)b * nvox * y_ldc + y_coff + sp.cg * 8;
    // NU voxels per thread per iteration: all loads are issued before the arithmetic
    const long long chunk = (long long)NU * sp.vstride;
    for (long long base = (long long)blockIdx.x * chunk; base < nvox; base += (long long)gridDim.x * chunk) {
      Raw8<T> d[NU], x[NU];
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const long long vv = min(base + sp.vlane + (long long)u * sp.vstride, nvox - 1);  // clamped tail re-reads
        d[u].load(dbase + vv * d_ldc);
        x[u].load(ybase + vv * y_ldc);
      }
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        if (base + sp.vlane + (long long)u * sp.vstride < nvox) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xv = x[u].get(j), dd = d[u].get(j);
            const float t = fmaf(xv, f[j].x, f[j].y);
            const float dv = t > 0.f ? dd : dd * f[j].z;
            const float xhat = (xv - mr[j].x) * mr[j].y;
            part[0][j] += dv;
            part[1][j] = fmaf(dv, xhat, part[1][j]);
          }
        }
      }
    }
  }
  reduce_store<2>(part, sp, C, red + (long long)b * C * 2);
}

int in_bwd_reduce(const void* dact, int d_ldc, int d_coff, const void* y, int y_ldc, int y_coff, int dtype, int B,
                  long long nvox, int C, const float* xform, const float* meanrstd, double* red, cudaStream_t s) {
  MTB_REQUIRE(C % 8 == 0 && C / 8 <= NT && d_ldc % 8 == 0 && d_coff % 8 == 0 && y_ldc % 8 == 0 && y_coff % 8 == 0,
              "in_bwd_reduce: channel counts/strides must be multiples of 8 (C=%d)", C);
  if (dtype != MTB200_F32) {
    const int r = in_bwd_reduce_tma(dact, d_ldc, d_coff, y, y_ldc, y_coff, dtype, B, nvox, C, xform, meanrstd, red, s);
    if (r != MTB200_ERR_UNSUPPORTED) return r;
  }
  dim3 grid = span_grid(nvox, B, C);
#define MTB_IN_RED(NU_)                                                                                          \
  MTB_DISPATCH_DTYPE(dtype, T, (in_bwd_reduce_kernel<T, NU_><<<grid, NT, 0, s>>>(                                \
      reinterpret_cast<const T*>(dact), d_ldc, d_coff, reinterpret_cast<const T*>(y), y_ldc, y_coff, nvox, C,    \
      reinterpret_cast<const float4*>(xform), reinterpret_cast<const float2*>(meanrstd), red)))
  switch (norm_nu()) { case 1: MTB_IN_RED(1); break; case 2: MTB_IN_RED(2); break; default: MTB_IN_RED(4); }
#undef MTB_IN_RED
  return check_launch("in_bwd_reduce");
}

// ---- backward pass 2 ------------------------------------------------------------------------------------------------
// `dy` may alias `dact` (in-place; how the engine runs it): those two pointers are therefore not __restrict__, and an
// aliased `dact` is read with coherent loads (ld.global, not the read-only .nc path; each thread reads its own element
// before writing it, the clamped tail re-reads are discarded).  A separate output buffer keeps both inputs on the
// read-only path, but measured SLOWER end to end (r2h: 5.03 vs 4.06 ms per step).
template <typename T, int NU>
__global__ void __launch_bounds__(NT) in_bwd_apply_kernel(const T* dact, int d_ldc, int d_coff,
                                                          const T* __restrict__ y, int y_ldc, int y_coff, T* dy,
                                                          int dy_ldc, int dy_coff, long long nvox, int B, int C,
                                                          const float4* __restrict__ xform,
                                                          const float2* __restrict__ meanrstd,
                                                          const float* __restrict__ gamma, const double* __restrict__ red,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const Span sp = make_span(nvox, C);
  const int b = blockIdx.y;
  // parameter gradients: sum over the batch, done once
  if (blockIdx.x == 0 && blockIdx.y == 0 && dgamma) {
    for (int c = threadIdx.x; c < C; c += NT) {
      double g = 0.0, bt = 0.0;
      for (int bb = 0; bb < B; ++bb) { bt += red[((long long)bb * C + c) * 2]; g += red[((long long)bb * C + c) * 2 + 1]; }
      dgamma[c] += (float)g;
      dbeta[c] += (float)bt;
    }
  }
  if (!sp.active) return;
  float4 f[8];
  float2 mr[8];
  float k1[8], m1[8], m2[8];
  const double inv_n = 1.0 / (double)nvox;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long i = (long long)b * C + sp.cg * 8 + j;
    f[j] = xform[i];
    mr[j] = meanrstd[i];
    k1[j] = mr[j].y * gamma[sp.cg * 8 + j];
    m1[j] = (float)(red[2 * i] * inv_n);
    m2[j] = (float)(red[2 * i + 1] * inv_n);
  }
  const T* dbase = dact + (long long)b * nvox * d_ldc + d_coff + sp.cg * 8;
  const T* ybase = y + (long long)b * nvox * y_ldc + y_coff + sp.cg * 8;
  T* obase = dy + (long long)b * nvox * dy_ldc + dy_coff + sp.cg * 8;
  const bool inplace = static_cast<const void*>(dact) == static_cast<const void*>(dy);  // uniform: read-only path unless aliased
  // NU voxels per thread per iteration: all loads are issued before the arithmetic
  const long long chunk = (long long)NU * sp.vstride;
  for (long long base = (long long)blockIdx.x * chunk; base < nvox; base += (long long)gridDim.x * chunk) {
    Raw8<T> d[NU], x[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const long long vv = min(base + sp.vlane + (long long)u * sp.vstride, nvox - 1);
      if (inplace) d[u].loadc(dbase + vv * d_ldc);  // dy aliases dact: coherent load
      else d[u].load(dbase + vv * d_ldc);
      x[u].load(ybase + vv * y_ldc);
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const long long vv = base + sp.vlane + (long long)u * sp.vstride;
      if (vv < nvox) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xv = x[u].get(j), dd = d[u].get(j);
          const float t = fmaf(xv, f[j].x, f[j].y);
          const float dv = t > 0.f ? dd : dd * f[j].z;
          const float xhat = (xv - mr[j].x) * mr[j].y;
          o[j] = k1[j] * (dv - m1[j] - xhat * m2[j]);
        }
        store8<T>(obase + vv * dy_ldc, o);
      }
    }
  }
}

// Second generation of the pass above: the per-channel constants live in SHARED memory (two float4 per channel) instead
// of 72 registers, so three to four blocks fit on an SM and ~4x more loads are in flight -- this pass mixes two read
// streams with one write stream and needs the extra memory-level parallelism to reach the HBM roofline.
//   dv = dact * lrelu'(sc * y + sh),   dy = k1 * (dv - m1 - xhat * m2) = k1 * dv + c1 * y + c0
template <typename T, int NU>
__global__ void __launch_bounds__(NT, 3) in_bwd_apply_smem_kernel(const T* dact, int d_ldc, int d_coff,
                                                                  const T* __restrict__ y, int y_ldc, int y_coff,
                                                                  T* dy, int dy_ldc, int dy_coff, long long nvox,
                                                                  int B, int C, const float4* __restrict__ xform,
                                                                  const float2* __restrict__ meanrstd,
                                                                  const float* __restrict__ gamma,
                                                                  const double* __restrict__ red, float* __restrict__ dgamma,
                                                                  float* __restrict__ dbeta) {
  extern __shared__ float4 s_const[];  // [C][2] = {sc, sh, slope, k1}, {c1, c0, -, -}
  const Span sp = make_span(nvox, C);
  const int b = blockIdx.y;
  if (blockIdx.x == 0 && blockIdx.y == 0 && dgamma) {
    for (int c = threadIdx.x; c < C; c += NT) {
      double g = 0.0, bt = 0.0;
      for (int bb = 0; bb < B; ++bb) { bt += red[((long long)bb * C + c) * 2]; g += red[((long long)bb * C + c) * 2 + 1]; }
      dgamma[c] += (float)g;
      dbeta[c] += (float)bt;
    }
  }
  const double inv_n = 1.0 / (double)nvox;
  for (int c = threadIdx.x; c < C; c += NT) {
    const long long i = (long long)b * C + c;
    const float4 f = xform[i];
    const float2 mr = meanrstd[i];
    const float k1 = mr.y * gamma[c];
    const float m1 = (float)(red[2 * i] * inv_n), m2 = (float)(red[2 * i + 1] * inv_n);
    s_const[2 * c] = make_float4(f.x, f.y, f.z, k1);
    s_const[2 * c + 1] = make_float4(-k1 * m2 * mr.y, k1 * (m2 * mr.y * mr.x - m1), 0.f, 0.f);
  }
  __syncthreads();
  if (!sp.active) return;
  const T* dbase = dact + (long long)b * nvox * d_ldc + d_coff + sp.cg * 8;
  const T* ybase = y + (long long)b * nvox * y_ldc + y_coff + sp.cg * 8;
  T* obase = dy + (long long)b * nvox * dy_ldc + dy_coff + sp.cg * 8;
  const bool inplace = static_cast<const void*>(dact) == static_cast<const void*>(dy);  // uniform: read-only path unless aliased
  const long long chunk = (long long)NU * sp.vstride;
  for (long long base = (long long)blockIdx.x * chunk; base < nvox; base += (long long)gridDim.x * chunk) {
    Raw8<T> d[NU], x[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const long long vv = min(base + sp.vlane + (long long)u * sp.vstride, nvox - 1);
      if (inplace) d[u].loadc(dbase + vv * d_ldc);  // dy aliases dact: coherent load
      else d[u].load(dbase + vv * d_ldc);
      x[u].load(ybase + vv * y_ldc);
    }
    int cg = sp.cg;
    asm volatile("" : "+r"(cg));  // keeps the constant fetches inside the loop (no 48-register hoist)
    const float4* cc = s_const + cg * 16;
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const long long vv = base + sp.vlane + (long long)u * sp.vstride;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 P = cc[2 * j], Q = cc[2 * j + 1];
        const float xv = x[u].get(j), dd = d[u].get(j);
        const float dv = fmaf(xv, P.x, P.y) > 0.f ? dd : dd * P.z;
        o[j] = fmaf(P.w, dv, fmaf(Q.x, xv, Q.y));
      }
      if (vv < nvox) store8<T>(obase + vv * dy_ldc, o);
    }
  }
}

int in_bwd_apply(const void* dact, int d_ldc, int d_coff, const void* y, int y_ldc, int y_coff, void* dy, int dy_ldc,
                 int dy_coff, int dtype, int B, long long nvox, int C, const float* xform, const float* meanrstd,
                 const float* gamma, const double* red, float* dgamma, float* dbeta, cudaStream_t s) {
  MTB_REQUIRE(C % 8 == 0 && C / 8 <= NT && d_ldc % 8 == 0 && d_coff % 8 == 0 && y_ldc % 8 == 0 && y_coff % 8 == 0 &&
                  dy_ldc % 8 == 0 && dy_coff % 8 == 0,
              "in_bwd_apply: channel counts/strides must be multiples of 8 (C=%d)", C);
  if (dtype != MTB200_F32) {  // TMA-pipelined version (norm_tma.cu); UNSUPPORTED = outside its envelope or switched off
    const int r = in_bwd_apply_tma(dact, d_ldc, d_coff, y, y_ldc, y_coff, dy, dy_ldc, dy_coff, dtype, B, nvox, C, xform,
                                   meanrstd, gamma, red, dgamma, dbeta, s);
    if (r != MTB200_ERR_UNSUPPORTED) return r;
  }
  dim3 grid = span_grid(nvox, B, C);
#define MTB_IN_APPLY(NU_)                                                                                        \
  MTB_DISPATCH_DTYPE(dtype, T, (in_bwd_apply_kernel<T, NU_><<<grid, NT, 0, s>>>(                                 \
      reinterpret_cast<const T*>(dact), d_ldc, d_coff, reinterpret_cast<const T*>(y), y_ldc, y_coff,             \
      reinterpret_cast<T*>(dy), dy_ldc, dy_coff, nvox, B, C, reinterpret_cast<const float4*>(xform),             \
      reinterpret_cast<const float2*>(meanrstd), gamma, red, dgamma, dbeta)))
#define MTB_IN_APPLY2(NU_)                                                                                       \
  MTB_DISPATCH_DTYPE(dtype, T, (in_bwd_apply_smem_kernel<T, NU_><<<grid, NT, (size_t)C * 32, s>>>(               \
      reinterpret_cast<const T*>(dact), d_ldc, d_coff, reinterpret_cast<const T*>(y), y_ldc, y_coff,             \
      reinterpret_cast<T*>(dy), dy_ldc, dy_coff, nvox, B, C, reinterpret_cast<const float4*>(xform),             \
      reinterpret_cast<const float2*>(meanrstd), gamma, red, dgamma, dbeta)))
  static const int variant = env_int("MTB200_APPLY_V", 2);
  if (variant == 2 && dtype != MTB200_F32) {
    switch (norm_nu()) { case 2: MTB_IN_APPLY2(2); break; case 8: MTB_IN_APPLY2(8); break; default: MTB_IN_APPLY2(4); }
  } else {
    switch (norm_nu()) { case 1: MTB_IN_APPLY(1); break; case 2: MTB_IN_APPLY(2); break; default: MTB_IN_APPLY(4); }
  }
#undef MTB_IN_APPLY
#undef MTB_IN_APPLY2
  return check_launch("in_bwd_apply");
}

// ---- plain LeakyReLU backward (second nonlinearity of a residual block) -----------------------------------------------
template <typename T>
__global__ void lrelu_bwd_kernel(const T* __restrict__ dact, const T* __restrict__ act, T* __restrict__ dv, long long n8,
                                 float slope) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float d[8], a[8];
    load8<T>(dact + i * 8, d);
    load8<T>(act + i * 8, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = a[j] > 0.f ? d[j] : d[j] * slope;
    store8<T>(dv + i * 8, d);
  }
}

int lrelu_bwd(const void* dact, const void* act, void* dv, int dtype, long long n, float slope, cudaStream_t s) {
  MTB_REQUIRE(n % 8 == 0, "lrelu_bwd: n must be a multiple of 8");
  const long long n8 = n / 8;
  const int blocks = (int)min((long long)num_sms() * 8, (n8 + 255) / 256);
  if (blocks == 0) return MTB200_OK;
  MTB_DISPATCH_DTYPE(dtype, T, (lrelu_bwd_kernel<T><<<blocks, 256, 0, s>>>(
      reinterpret_cast<const T*>(dact), reinterpret_cast<const T*>(act), reinterpret_cast<T*>(dv), n8, slope)));
  return check_launch("lrelu_bwd");
}

// ---- tail of a residual block, backward: dv = dact * lrelu'(act) fanned out to the two summands -----------------------
// out = lrelu(f(a) + g(r))  =>  d f(a) = d g(r) = dact * (out > 0 ? 1 : slope); each destination is written or accumulated.
template <typename T>
__global__ void __launch_bounds__(NT) residual_bwd_kernel(const T* __restrict__ dact, int d_ldc, int d_coff,
                                                          const T* __restrict__ act, int a_ldc, int a_coff, T* dst0,
                                                          int ldc0, int coff0, int acc0, T* dst1, int ldc1, int coff1,
                                                          int acc1, long long nrows, int C, float slope) {
  const int G = C / 8;
  const long long total = nrows * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / G;
    const int c = (int)(i % G) * 8;
    float d[8], a[8];
    load8<T>(dact + v * d_ldc + d_coff + c, d);
    load8<T>(act + v * a_ldc + a_coff + c, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = a[j] > 0.f ? d[j] : d[j] * slope;
    if (dst0) {
      float o[8];
      T* q = dst0 + v * ldc0 + coff0 + c;
      if (acc0) {
        load8<T>(q, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += d[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = d[j];
      }
      store8<T>(q, o);
    }
    if (dst1) {
      float o[8];
      T* q = dst1 + v * ldc1 + coff1 + c;
      if (acc1) {
        load8<T>(q, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += d[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = d[j];
      }
      store8<T>(q, o);
    }
  }
}

int residual_bwd(const void* dact, int d_ldc, int d_coff, const void* act, int a_ldc, int a_coff, void* dst0, int ldc0,
                 int coff0, int acc0, void* dst1, int ldc1, int coff1, int acc1, int dtype, long long nrows, int C,
                 float slope, cudaStream_t s) {
  MTB_REQUIRE(C % 8 == 0 && d_ldc % 8 == 0 && d_coff % 8 == 0 && a_ldc % 8 == 0 && a_coff % 8 == 0 && ldc0 % 8 == 0 &&
                  coff0 % 8 == 0 && ldc1 % 8 == 0 && coff1 % 8 == 0,
              "residual_bwd: channel counts/strides must be multiples of 8 (C=%d)", C);
  const long long total = nrows * (C / 8);
  if (total == 0) return MTB200_OK;
  const int blocks = (int)min((long long)num_sms() * 16, (total + NT - 1) / NT);
  MTB_DISPATCH_DTYPE(dtype, T, (residual_bwd_kernel<T><<<blocks, NT, 0, s>>>(
      reinterpret_cast<const T*>(dact), d_ldc, d_coff, reinterpret_cast<const T*>(act), a_ldc, a_coff,
      reinterpret_cast<T*>(dst0), ldc0, coff0, acc0, reinterpret_cast<T*>(dst1), ldc1, coff1, acc1, nrows, C, slope)));
  return check_launch("residual_bwd");
}

}  // namespace mtb
