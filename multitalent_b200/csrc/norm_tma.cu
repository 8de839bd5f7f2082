// TMA-pipelined versions of the InstanceNorm streaming passes (16-bit tensors, sm_100a).
//
// norm.cu's passes keep their loads in flight from registers: every warp drains its loads before it computes, so the
// bytes in flight swing between "all" and "none" and the 2-reads-1-write pass (in_bwd_apply) stops at ~4.1 TB/s while a
// plain elementwise add reaches 6.9 TB/s on the same tensors (tools/bw_probe.py).  Here the memory side is decoupled
// from the arithmetic: one producer warp keeps a ring of 3 stages of bulk tensor loads (2 x 16 KB each) permanently in
// flight, 8 consumer warps transform shared-memory tiles (each warp its own 2 KB slice, per-channel constants in
// registers, no CTA-wide barrier), and every warp sends its slice back with a bulk tensor store -- the SM always has
// ~100 KB outstanding regardless of what the consumers are doing.
//
//   MODE 0  in_bwd_apply:  dy = k1 * dv + c1 * y + c0,  dv = dact * lrelu'(sc * y + sh)   (InstanceNorm + LeakyReLU backward,
//           generic_UNet.py:63-70; constants per (b, c) from the reduction pass), in place or out of place
//   MODE 1  norm_act:      act = lrelu(sc * y + sh)                                        (materialise)
//   MODE 2  in_bwd_reduce: red[b][c] += {sum dv, sum dv * xhat}  (no output tile; per-lane partial sums in registers for
//           the whole CTA, combined in a fixed order inside the CTA, fp64 atomics across CTAs)
//
// A tile = R voxels x C channels of one sample (3-D tensor maps {C, nvox, B}: the ragged last tile of a sample is zero
// filled on load and clipped on store); every CTA works on ONE sample, so the per-channel constants sit in shared memory
// once.  The arithmetic is norm.cu's, operation for operation: results are bit-identical.
#include "umma.cuh"

namespace mtb {

using namespace um;

constexpr int ST_CONSUMERS = 256;
constexpr int ST_THREADS = ST_CONSUMERS + 32;
constexpr int ST_STAGES = 3;

struct StreamParams {
  CUtensorMap a_map, b_map, o_map;  // MODE 0: a = dact, b = y, o = dy;  MODE 1: a = y, o = act
  const float4* xform;
  const float2* meanrstd;
  const float* gamma;
  const double* red;
  float* dgamma;
  float* dbeta;
  double* red_out;  // MODE 2
  long long nvox;
  int B, C, R, nbox, cbox, cps, tiles_per_b, tile_bytes, box_bytes;
};

__device__ __forceinline__ void st_tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

template <typename T> struct StVec;
template <> struct StVec<__nv_bfloat16> {
  static __device__ __forceinline__ float lo(uint32_t w) { return __uint_as_float(w << 16); }
  static __device__ __forceinline__ float hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};
template <> struct StVec<__half> {
  static __device__ __forceinline__ float lo(uint32_t w) { return __low2float(*reinterpret_cast<const __half2*>(&w)); }
  static __device__ __forceinline__ float hi(uint32_t w) { return __high2float(*reinterpret_cast<const __half2*>(&w)); }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

template <typename T, int MODE>
__global__ void __launch_bounds__(ST_THREADS, 1) in_stream_tma_kernel(const __grid_constant__ StreamParams p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  constexpr int NIN = MODE == 1 ? 1 : 2;
  constexpr int NOUT = MODE == 2 ? 0 : 1;
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t full[ST_STAGES], empty[ST_STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = (NIN + NOUT) * p.tile_bytes;
  float4* s_const = reinterpret_cast<float4*>(dsmem + ST_STAGES * stage_bytes);  // [C][2]
  const int b = (int)blockIdx.x / p.cps, slot = (int)blockIdx.x % p.cps;

  if (threadIdx.x == 0) {
    for (int i = 0; i < ST_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], ST_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (MODE == 0) {
    if (blockIdx.x == 0 && p.dgamma) {  // parameter gradients: sum over the batch, done once
      for (int c = threadIdx.x; c < p.C; c += ST_THREADS) {
        double g = 0.0, bt = 0.0;
        for (int bb = 0; bb < p.B; ++bb) { bt += p.red[((long long)bb * p.C + c) * 2]; g += p.red[((long long)bb * p.C + c) * 2 + 1]; }
        p.dgamma[c] += (float)g;
        p.dbeta[c] += (float)bt;
      }
    }
    const double inv_n = 1.0 / (double)p.nvox;
    for (int c = threadIdx.x; c < p.C; c += ST_THREADS) {
      const long long i = (long long)b * p.C + c;
      const float4 f = p.xform[i];
      const float2 mr = p.meanrstd[i];
      const float k1 = mr.y * p.gamma[c];
      const float m1 = (float)(p.red[2 * i] * inv_n), m2 = (float)(p.red[2 * i + 1] * inv_n);
      s_const[2 * c] = make_float4(f.x, f.y, f.z, k1);
      s_const[2 * c + 1] = make_float4(-k1 * m2 * mr.y, k1 * (m2 * mr.y * mr.x - m1), 0.f, 0.f);
    }
  } else {
    for (int c = threadIdx.x; c < p.C; c += ST_THREADS) {
      s_const[2 * c] = p.xform ? p.xform[(long long)b * p.C + c] : make_float4(1.f, 0.f, 1.f, 0.f);
      float2 mr = make_float2(0.f, 0.f);
      if (MODE == 2) mr = p.meanrstd[(long long)b * p.C + c];
      s_const[2 * c + 1] = make_float4(mr.x, mr.y, 0.f, 0.f);
    }
  }
  __syncthreads();

  if (warp == ST_CONSUMERS / 32) {
    // ===== producer =====
    uint32_t k = 0;
    for (int t = slot; t < p.tiles_per_b; t += p.cps, ++k) {
      const uint32_t s = k % ST_STAGES;
      mbar_wait(&empty[s], ((k / ST_STAGES) & 1u) ^ 1u);
      if (elect_one()) {
        uint8_t* dst = dsmem + (size_t)s * stage_bytes;
        mbar_expect_tx(&full[s], (uint32_t)(NIN * p.tile_bytes));
        for (int j = 0; j < p.nbox; ++j) {
          tma_load_3d(dst + j * p.box_bytes, &p.a_map, &full[s], j * p.cbox, t * p.R, b);
          if (NIN == 2) tma_load_3d(dst + p.tile_bytes + j * p.box_bytes, &p.b_map, &full[s], j * p.cbox, t * p.R, b);
        }
      }
      __syncwarp();
    }
  } else {
    // ===== consumers: every warp owns a 2 KB slice (128 vectors) of each tile and stores it itself -- no CTA barrier =====
    const int Gb = p.cbox / 8;           // 16-byte vectors per row; 4 / 8 / 16 / 32, so a lane always sees one channel group
    const int cg = lane % Gb;
    float4 P[8];
    float2 Q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      P[j] = s_const[2 * (cg * 8 + j)];
      const float4 qq = s_const[2 * (cg * 8 + j) + 1];
      Q[j] = make_float2(qq.x, qq.y);
    }
    const int rows_w = p.R / (ST_CONSUMERS / 32);  // rows of a tile per warp
    float acc0[8], acc1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc0[j] = acc1[j] = 0.f;
    uint32_t k = 0;
    for (int t = slot; t < p.tiles_per_b; t += p.cps, ++k) {
      const uint32_t s = k % ST_STAGES;
      const uint4* in0 = reinterpret_cast<const uint4*>(dsmem + (size_t)s * stage_bytes) + warp * 128;
      const uint4* in1 = reinterpret_cast<const uint4*>(dsmem + (size_t)s * stage_bytes + p.tile_bytes) + warp * 128;
      uint4* out = reinterpret_cast<uint4*>(dsmem + (size_t)s * stage_bytes + NIN * p.tile_bytes) + warp * 128;
      // the store this warp issued from the same output slice three tiles ago must have read its shared memory
      if (NOUT) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(ST_STAGES - 1) : "memory");
        __syncwarp();
      }
      mbar_wait(&full[s], (k / ST_STAGES) & 1u);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int i = it * 32 + lane;
        const uint4 a = in0[i];
        uint4 x = a;
        if (NIN == 2) x = in1[i];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, xw[4] = {x.x, x.y, x.z, x.w};
        uint32_t ow[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float r[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 PP = P[2 * e + h];
            const float xv = h ? StVec<T>::hi(xw[e]) : StVec<T>::lo(xw[e]);
            if (MODE == 0) {
              const float dd = h ? StVec<T>::hi(aw[e]) : StVec<T>::lo(aw[e]);
              const float dv = fmaf(xv, PP.x, PP.y) > 0.f ? dd : dd * PP.z;
              r[h] = fmaf(PP.w, dv, fmaf(Q[2 * e + h].x, xv, Q[2 * e + h].y));
            } else if (MODE == 2) {
              const float dd = h ? StVec<T>::hi(aw[e]) : StVec<T>::lo(aw[e]);
              const float dv = fmaf(xv, PP.x, PP.y) > 0.f ? dd : dd * PP.z;
              const float xhat = (xv - Q[2 * e + h].x) * Q[2 * e + h].y;
              acc0[2 * e + h] += dv;
              acc1[2 * e + h] = fmaf(dv, xhat, acc1[2 * e + h]);
              r[h] = 0.f;
            } else {
              const float tt = fmaf(xv, PP.x, PP.y);
              r[h] = tt > 0.f ? tt : tt * PP.z;
            }
          }
          if (NOUT) ow[e] = StVec<T>::pack(r[0], r[1]);
        }
        if (NOUT) out[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
      if (NOUT) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&empty[s]);  // this warp is done with the input tiles of the stage
        if (NOUT) {
          st_tma_store_3d(&p.o_map, out, 0, t * p.R + warp * rows_w, b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (NOUT && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (MODE == 2) {
      // lanes of one channel group (lane % Gb), then the 8 warps, in a fixed order; one fp64 atomic per (channel, sum)
      for (int off = 16; off >= Gb; off >>= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc0[j] += __shfl_xor_sync(0xffffffffu, acc0[j], off);
          acc1[j] += __shfl_xor_sync(0xffffffffu, acc1[j], off);
        }
      }
      float* s_part = reinterpret_cast<float*>(dsmem);  // [warp][C][2]: the stage ring is idle now
      asm volatile("bar.sync 1, %0;" ::"n"(ST_CONSUMERS) : "memory");  // every consumer warp has left its last tile
      if (lane < Gb) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s_part[(warp * p.C + cg * 8 + j) * 2] = acc0[j];
          s_part[(warp * p.C + cg * 8 + j) * 2 + 1] = acc1[j];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(ST_CONSUMERS) : "memory");
      for (int idx = threadIdx.x; idx < p.C * 2; idx += ST_CONSUMERS) {
        float v = 0.f;
        for (int w = 0; w < ST_CONSUMERS / 32; ++w) v += s_part[w * p.C * 2 + idx];
        if (v != 0.f) atomicAdd(p.red_out + (long long)b * p.C * 2 + idx, (double)v);
      }
    }
  }
}

int umma_available();

static bool st_env_on(const char* name) {
  const char* e = getenv(name);
  return !e || atoi(e) != 0;
}

// Returns MTB200_ERR_UNSUPPORTED when the problem is outside the envelope (the caller runs the register-path kernel).
static int stream_tma_launch(int mode, const void* a, int a_ldc, int a_coff, const void* bsrc, int b_ldc, int b_coff, void* o,
                             int o_ldc, int o_coff, int dtype, int B, long long nvox, int C, const float* xform,
                             const float* meanrstd, const float* gamma, const double* red, float* dgamma, float* dbeta,
                             double* red_out, cudaStream_t s) {
  if (!umma_available() || (dtype != MTB200_BF16 && dtype != MTB200_F16)) return MTB200_ERR_UNSUPPORTED;
  if (C % 8 || nvox >= (1LL << 31) || B > 65535) return MTB200_ERR_UNSUPPORTED;
  // 16 KB tiles of R rows, one box; every consumer warp owns R / 8 rows = 128 vectors whose channel group depends on the
  // lane only: C = 32 / 64 / 128 / 256 (the 320-channel levels are a few hundred KB: they keep the register-path kernel)
  if (C != 32 && C != 64 && C != 128 && C != 256) return MTB200_ERR_UNSUPPORTED;
  const int nbox = 1, cbox = C;
  const int R = 16384 / (C * 2);
  if (nvox < 4LL * R) return MTB200_ERR_UNSUPPORTED;
  static thread_local StreamParams q;
  memset(&q, 0, sizeof(q));
  auto mk = [&](CUtensorMap* m, const void* base, int ldc, int coff, int rows) {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)nvox, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ldc * 2, (cuuint64_t)nvox * ldc * 2};
    cuuint32_t box[3] = {(cuuint32_t)cbox, (cuuint32_t)rows, 1};
    return umma_encode_map(m, dtype, 3, (uint8_t*)base + (size_t)coff * 2, dims, strides, box, 0);
  };
  if (!mk(&q.a_map, a, a_ldc, a_coff, R)) return MTB200_ERR_CUDA;
  if (mode != 1 && !mk(&q.b_map, bsrc, b_ldc, b_coff, R)) return MTB200_ERR_CUDA;
  if (mode != 2 && !mk(&q.o_map, o, o_ldc, o_coff, R / (ST_CONSUMERS / 32))) return MTB200_ERR_CUDA;  // one warp's slice per store
  q.xform = reinterpret_cast<const float4*>(xform); q.meanrstd = reinterpret_cast<const float2*>(meanrstd);
  q.gamma = gamma; q.red = red; q.dgamma = dgamma; q.dbeta = dbeta; q.red_out = red_out;
  q.nvox = nvox; q.B = B; q.C = C; q.R = R; q.nbox = nbox; q.cbox = cbox;
  q.tiles_per_b = (int)((nvox + R - 1) / R);
  q.box_bytes = R * cbox * 2;
  q.tile_bytes = q.box_bytes * nbox;
  int cps = num_sms() / B;
  if (cps < 1) cps = 1;
  if (cps > q.tiles_per_b) cps = q.tiles_per_b;
  q.cps = cps;
  const int ntile = mode == 1 ? 2 : (mode == 0 ? 3 : 2);  // tiles per stage (inputs + output)
  const int smem = ST_STAGES * ntile * q.tile_bytes + C * 32 + 1024;
  dim3 grid((unsigned)(cps * B));
  cudaError_t e = cudaSuccess;
#define ST_LAUNCH(T, MODE)                                                                                         \
  do {                                                                                                             \
    e = cudaFuncSetAttribute(in_stream_tma_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);    \
    if (e == cudaSuccess) launch_pdl(in_stream_tma_kernel<T, MODE>, dim3(grid), dim3(ST_THREADS), (size_t)(smem), s, q);                         \
  } while (0)
  if (dtype == MTB200_BF16) {
    if (mode == 0) ST_LAUNCH(__nv_bfloat16, 0); else if (mode == 1) ST_LAUNCH(__nv_bfloat16, 1); else ST_LAUNCH(__nv_bfloat16, 2);
  } else {
    if (mode == 0) ST_LAUNCH(__half, 0); else if (mode == 1) ST_LAUNCH(__half, 1); else ST_LAUNCH(__half, 2);
  }
#undef ST_LAUNCH
  if (e != cudaSuccess) { set_error("in_stream_tma: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  return MTB200_OK;
}

int in_bwd_apply_tma(const void* dact, int d_ldc, int d_coff, const void* y, int y_ldc, int y_coff, void* dy, int dy_ldc,
                     int dy_coff, int dtype, int B, long long nvox, int C, const float* xform, const float* meanrstd,
                     const float* gamma, const double* red, float* dgamma, float* dbeta, cudaStream_t s) {
  static const bool on = st_env_on("MTB200_APPLY_TMA");
  if (!on) return MTB200_ERR_UNSUPPORTED;
  const int r = stream_tma_launch(0, dact, d_ldc, d_coff, y, y_ldc, y_coff, dy, dy_ldc, dy_coff, dtype, B, nvox, C, xform,
                                  meanrstd, gamma, red, dgamma, dbeta, nullptr, s);
  return r == MTB200_OK ? check_launch("in_bwd_apply_tma") : r;
}

int norm_act_tma(const void* y, int in_ldc, int in_coff, void* out, int out_ldc, int out_coff, int dtype, int B,
                 long long nvox, int C, const float* xform, cudaStream_t s) {
  static const bool on = st_env_on("MTB200_NORM_TMA");
  if (!on) return MTB200_ERR_UNSUPPORTED;
  const int r = stream_tma_launch(1, y, in_ldc, in_coff, nullptr, 0, 0, out, out_ldc, out_coff, dtype, B, nvox, C, xform,
                                  nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, s);
  return r == MTB200_OK ? check_launch("norm_act_tma") : r;
}

int in_bwd_reduce_tma(const void* dact, int d_ldc, int d_coff, const void* y, int y_ldc, int y_coff, int dtype, int B,
                      long long nvox, int C, const float* xform, const float* meanrstd, double* red, cudaStream_t s) {
  static const bool on = st_env_on("MTB200_REDUCE_TMA");
  if (!on) return MTB200_ERR_UNSUPPORTED;
  const int r = stream_tma_launch(2, dact, d_ldc, d_coff, y, y_ldc, y_coff, nullptr, 0, 0, dtype, B, nvox, C, xform, meanrstd,
                                  nullptr, nullptr, nullptr, nullptr, red, s);
  return r == MTB200_OK ? check_launch("in_bwd_reduce_tma") : r;
}

}  // namespace mtb
