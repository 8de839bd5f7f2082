// Line-streaming tcgen05 weight gradient of the STRIDED 3x3x3 convolutions (stride 2 in every axis, padding 1; the first
// conv of every encoder stage, generic_UNet.py:126-128 with convolutional pooling):
//     dW[(dz,dy,dx)][co][ci] = sum_{b, o} dY[b, o][co] * X[b, 2 o + (dz,dy,dx)][ci]
// The per-tap kernel fetched one strided 64-byte-row brick per (tap, brick) -- 27 re-fetches of X through the L2 -> SM
// path -- and ran at 160 TFLOP/s on the 32 -> 64 layer (1.19 ms for 0.2 ms of HBM traffic).  Here, as in wgrad_line.cu, the
// GEMM K dimension is the voxels of one OUTPUT line (16 per MMA) and both operands are MN-major straight out of NDHWC:
//   * an input line h' of plane d' is staged as THREE parity boxes by TMA element strides (traversal stride 2 along w):
//     O1 = x[2i - 1], E = x[2i], O2 = x[2i + 1], i = 0 .. wt-1 -- the dx = -1 / 0 / +1 operands of output voxel i.  They sit
//     one region apart, so ONE descriptor with LBO = region size presents them as the M blocks of A (32 ci each; the
//     fourth block reads whatever follows and its rows are discarded);
//   * an ODD input line h' = 2 oh + 1 = 2 (oh + 1) - 1 meets the output lines oh (dy = +1) and oh + 1 (dy = -1): B = two dY
//     lines side by side (N = 64); an EVEN line h' = 2 oh meets only oh (dy = 0, N = 32).  Lines of another work unit's
//     range (or outside the volume) are fetched from an out-of-bounds coordinate = zeros;
//   * planes d' = 2 od - 1, 2 od, 2 od + 1 (dz) are three boxes sets per step, all against plane od of dY;
//   * accumulators resident in TMEM for the whole kernel: per dz one [128 x 64] (odd lines) and one [128 x 32] (even lines).
// Persistent CTAs, the (Cin chunk, Cout block) pairs of a unit co-scheduled on neighbouring CTAs, fp32 atomics once per CTA.
#include "umma.cuh"

namespace mtb {

using namespace um;

constexpr int WS_THREADS = 192;
constexpr int WS_MAX_STAGES = 6;
constexpr int WS_BN = 32;   // Cout block
constexpr int WS_KC = 32;   // Cin chunk (64-byte rows)

struct WgradS2Params {
  CUtensorMap x_map, dy_map;   // x: element stride 2 along w
  float* dw;
  int B, Do, Ho, Wo;           // output grid = dY extent
  int Hi;                      // input lines per plane
  int Cin, Cout;
  int wt, nkk;                 // output voxels per line tile, 16-voxel K steps
  int region_bytes;            // one parity box, 1024-aligned
  int plane_bytes;             // 3 regions (+ slack for the discarded fourth block)
  int ybase, yline_bytes;
  int stage_bytes, stages;
  int x_tx, y_tx;              // bytes per parity box / per dY line
  int nhr, hlen, ntw;
  long long units;
  int npy, npz, nslots;
  int lut[27];
  int is_f16;
};

__device__ __forceinline__ uint64_t ws_desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | (uint64_t)lo; }

__global__ void __launch_bounds__(WS_THREADS, 1) wgrad_line_s2_umma_kernel(const __grid_constant__ WgradS2Params p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t st_full[WS_MAX_STAGES], st_empty[WS_MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t tmem_slot;

  constexpr uint32_t ROWB = WS_KC * 2;
  constexpr uint32_t ACC_ODD = 64, ACC_EVEN = 32, ACC_DZ = ACC_ODD + ACC_EVEN;  // TMEM columns per dz
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  const int npairs = p.npy * p.npz;
  const int pair = (int)blockIdx.x % npairs, slot = (int)blockIdx.x / npairs;
  const int c0 = (pair % p.npy) * WS_KC;
  const int n0 = (pair / p.npy) * WS_BN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], 1); }
    mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const bool have_work = (long long)slot < p.units;

  // unit -> (b, od, oh range, w tile); its input lines h' = 2 ohs - 1 .. 2 ohe - 1 (clipped at 0)
  auto decode = [&](long long u, int& b, int& od, int& ohs, int& ohe, int& ow0, int& h0, int& h1) {
    const int hr = (int)(u % p.nhr); u /= p.nhr;
    const int twi = (int)(u % p.ntw); u /= p.ntw;
    od = (int)(u % p.Do);
    b = (int)(u / p.Do);
    ohs = hr * p.hlen;
    ohe = min(p.Ho, ohs + p.hlen);
    ow0 = twi * p.wt;
    h0 = max(0, 2 * ohs - 1);
    h1 = min(p.Hi - 1, 2 * ohe - 1);
  };

  if (warp == 0) {
    // ===== producer =====
    uint32_t sc = 0;
    for (long long u = slot; u < p.units; u += p.nslots) {
      int b, od, ohs, ohe, ow0, h0, h1;
      decode(u, b, od, ohs, ohe, ow0, h0, h1);
      for (int hp = h0; hp <= h1; ++hp, ++sc) {
        const uint32_t slot_s = sc % (uint32_t)p.stages;
        mbar_wait(&st_empty[slot_s], ((sc / (uint32_t)p.stages) & 1u) ^ 1u);
        if (elect_one()) {
          uint8_t* dst = dsmem + (size_t)slot_s * p.stage_bytes;
          const bool odd = hp & 1;
          mbar_expect_tx(&st_full[slot_s], (uint32_t)(9 * p.x_tx + (odd ? 2 : 1) * p.y_tx));
          for (int z = 0; z < 3; ++z) {
            uint8_t* pl = dst + (size_t)z * p.plane_bytes;
            const int dp = 2 * od - 1 + z;
            tma_load_5d(pl, &p.x_map, &st_full[slot_s], c0, 2 * ow0 - 1, hp, dp, b);                       // O1: dx = -1
            tma_load_5d(pl + p.region_bytes, &p.x_map, &st_full[slot_s], c0, 2 * ow0, hp, dp, b);          // E:  dx =  0
            tma_load_5d(pl + 2 * p.region_bytes, &p.x_map, &st_full[slot_s], c0, 2 * ow0 + 1, hp, dp, b);  // O2: dx = +1
          }
          const int oob = -8;  // any out-of-bounds line: zero fill
          if (odd) {
            const int oa = (hp - 1) >> 1, ob = (hp + 1) >> 1;  // dy = +1, dy = -1
            tma_load_5d(dst + p.ybase, &p.dy_map, &st_full[slot_s], n0, ow0, (oa >= ohs && oa < ohe) ? oa : oob, od, b);
            tma_load_5d(dst + p.ybase + p.yline_bytes, &p.dy_map, &st_full[slot_s], n0, ow0,
                        (ob >= ohs && ob < ohe) ? ob : oob, od, b);
          } else {
            tma_load_5d(dst + p.ybase, &p.dy_map, &st_full[slot_s], n0, ow0, hp >> 1, od, b);               // dy = 0
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t fmt = p.is_f16 ? 0u : 1u;
    const uint32_t idesc0 = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((128u >> 4) << 24);
    const uint32_t idesc_odd = idesc0 | ((ACC_ODD >> 3) << 17), idesc_even = idesc0 | ((ACC_EVEN >> 3) << 17);
    const uint32_t hi_a = ((8u * ROWB) >> 4) | (1u << 14) | (4u << 29);  // SWIZZLE_64B, SBO = 8 rows
    const uint32_t lbo_a = ((uint32_t)p.region_bytes >> 4) << 16;       // M blocks = the three parity boxes
    const uint32_t hi_b = hi_a;
    const uint32_t lbo_b = ((uint32_t)p.yline_bytes >> 4) << 16;        // N blocks = the dY lines
    const uint32_t s16 = __shfl_sync(0xffffffffu, (smem_u32(dsmem) & 0x3FFFFu) >> 4, 0);
    const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, plane16 = (uint32_t)p.plane_bytes >> 4;
    const uint32_t y16 = (uint32_t)p.ybase >> 4;
    uint32_t sc = 0;
    uint32_t seen = 0;  // bit 0: an even line was accumulated, bit 1: an odd one
    for (long long u = slot; u < p.units; u += p.nslots) {
      int b, od, ohs, ohe, ow0, h0, h1;
      decode(u, b, od, ohs, ohe, ow0, h0, h1);
      for (int hp = h0; hp <= h1; ++hp, ++sc) {
        const uint32_t slot_s = sc % (uint32_t)p.stages;
        mbar_wait(&st_full[slot_s], (sc / (uint32_t)p.stages) & 1u);
        tc_fence_after();
        const uint32_t odd = (uint32_t)(hp & 1);
        if (elect_one()) {
          const uint32_t a_s = (s16 + slot_s * stage16) | lbo_a;
          const uint32_t b_s = (s16 + slot_s * stage16 + y16) | lbo_b;
          const uint32_t acc = (seen >> odd) & 1u;
          const uint32_t idesc = odd ? idesc_odd : idesc_even;
          const uint32_t col0 = odd ? 0u : ACC_ODD;
          // kk-major: consecutive MMAs go to different accumulators (dz)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kk < p.nkk) {
#pragma unroll
              for (int z = 0; z < 3; ++z)
                umma_f16(tmem_u + (uint32_t)z * ACC_DZ + col0, ws_desc64(hi_a, a_s + (uint32_t)z * plane16 + (uint32_t)(kk * ROWB)),
                         ws_desc64(hi_b, b_s + (uint32_t)(kk * ROWB)), idesc, kk ? 1u : acc);
            }
          }
          umma_commit(&st_empty[slot_s]);
        }
        __syncwarp();
        seen |= 1u << odd;
      }
    }
    if (have_work) {
      if (elect_one()) umma_commit(&acc_full);
      __syncwarp();
    }
  } else if (have_work) {
    // ===== epilogue: TMEM -> fp32 atomics into dW[widx][co][ci] =====
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int j = m / WS_KC, ci = m % WS_KC;  // j = dx + 1 (3 = the discarded block)
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    for (int z = 0; z < 3; ++z) {
      for (int c16 = 0; c16 < (int)ACC_DZ / 16; ++c16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)z * ACC_DZ + (uint32_t)(c16 * 16), r);
        const int col = c16 * 16;
        // columns 0..31: dy = +1, 32..63: dy = -1 (odd lines); 64..95: dy = 0 (even lines)
        const int dyi = col < 32 ? 2 : (col < 64 ? 0 : 1);
        const int co = col % 32;
        const int widx = j <= 2 ? p.lut[(z * 3 + dyi) * 3 + j] : -1;
        if (widx >= 0) {
          float* dst = p.dw + ((long long)widx * p.Cout + n0 + co) * p.Cin + c0 + ci;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float v = __uint_as_float(r[e]);
            if (v != 0.f) atomicAdd(dst + (long long)e * p.Cin, v);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

static inline int ws_align1k(long long v) { return (int)(((v + 1023) / 1024) * 1024); }

// Returns MTB200_ERR_UNSUPPORTED when the problem is outside this kernel's envelope (caller falls back).
int wgrad_line_s2_umma(const mtb200_wgrad_params& p, cudaStream_t s) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("MTB200_WLINE_S2"); enabled = (e && atoi(e) == 0) ? 0 : 1; }
  if (!enabled || p.ngroups != 1 || p.xform || p.ntaps != 27) return MTB200_ERR_UNSUPPORTED;
  for (int k = 0; k < 3; ++k)
    if (p.is[k] != 2 || p.os[k] != 1 || p.group_ooff[0][k] != 0) return MTB200_ERR_UNSUPPORTED;
  if (p.Di != 2 * p.Do || p.Hi != 2 * p.Ho || p.Wi != 2 * p.Wo || p.Dof != p.Do || p.Hof != p.Ho || p.Wof != p.Wo)
    return MTB200_ERR_UNSUPPORTED;
  // 32-voxel output lines (64 -> 128 at 48x40x32) measured 0.357 ms here against 0.313 ms on the per-tap kernel: two K
  // steps per line do not amortise nine parity boxes; MTB200_WLINE_S2_MINW overrides the threshold
  static const int minw = [] { const char* e = getenv("MTB200_WLINE_S2_MINW"); return e ? atoi(e) : 64; }();
  if (p.Cin % WS_KC || p.Cout % WS_BN || p.Wo < 32 || p.Wo < minw || p.Ho < 4) return MTB200_ERR_UNSUPPORTED;
  const int npy = p.Cin / WS_KC, npz = p.Cout / WS_BN;
  if (npy * npz > 16) return MTB200_ERR_UNSUPPORTED;

  static thread_local WgradS2Params q;
  memset(&q, 0, sizeof(q));
  for (int i = 0; i < 27; ++i) q.lut[i] = -1;
  for (int t = 0; t < p.ntaps; ++t) {
    for (int k = 0; k < 3; ++k)
      if (p.tap_off[t][k] < -1 || p.tap_off[t][k] > 1) return MTB200_ERR_UNSUPPORTED;
    int& e = q.lut[((p.tap_off[t][0] + 1) * 3 + (p.tap_off[t][1] + 1)) * 3 + (p.tap_off[t][2] + 1)];
    if (e >= 0) return MTB200_ERR_UNSUPPORTED;
    e = p.tap_widx[t];
  }
  q.wt = p.Wo > 32 ? 64 : 32;
  q.nkk = q.wt / 16;
  const int rows = q.wt;                                      // box r-th row = x[start + 2 r]: all three parities share the K index
  q.x_tx = rows * WS_KC * 2;
  q.region_bytes = ws_align1k(q.x_tx);
  q.plane_bytes = 4 * q.region_bytes;                         // the discarded fourth M block stays inside the stage
  q.yline_bytes = ws_align1k((long long)q.wt * WS_BN * 2);
  q.y_tx = q.wt * WS_BN * 2;
  q.ybase = 3 * q.plane_bytes;
  q.stage_bytes = q.ybase + 2 * q.yline_bytes;
  q.stages = min(WS_MAX_STAGES, (224 * 1024) / q.stage_bytes);
  if (q.stages < 2) return MTB200_ERR_UNSUPPORTED;
  {
    EncodeTiledFn enc = umma_encode_fn();
    if (!enc) return MTB200_ERR_UNSUPPORTED;
    const CUtensorMapDataType dt = p.dtype == MTB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.Di, (cuuint64_t)p.B};
    cuuint64_t strides[4] = {(cuuint64_t)p.in_ldc * 2, (cuuint64_t)p.Wi * p.in_ldc * 2,
                             (cuuint64_t)p.Hi * p.Wi * p.in_ldc * 2, (cuuint64_t)p.Di * p.Hi * p.Wi * p.in_ldc * 2};
    cuuint32_t box[5] = {(cuuint32_t)WS_KC, (cuuint32_t)(2 * rows - 1), 1, 1, 1};  // traversal extent: `rows` elements at stride 2
    cuuint32_t estr[5] = {1, 2, 1, 1, 1};
    CUresult r = enc(&q.x_map, dt, 5, (uint8_t*)p.x + (size_t)p.in_coff * 2, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { enabled = 0; return MTB200_ERR_UNSUPPORTED; }  // driver refuses element strides here
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.Cout, (cuuint64_t)p.Wof, (cuuint64_t)p.Hof, (cuuint64_t)p.Dof, (cuuint64_t)p.B};
    cuuint64_t strides[4] = {(cuuint64_t)p.out_ldc * 2, (cuuint64_t)p.Wof * p.out_ldc * 2,
                             (cuuint64_t)p.Hof * p.Wof * p.out_ldc * 2,
                             (cuuint64_t)p.Dof * p.Hof * p.Wof * p.out_ldc * 2};
    cuuint32_t box[5] = {(cuuint32_t)WS_BN, (cuuint32_t)q.wt, 1, 1, 1};
    if (!umma_encode_map(&q.dy_map, p.dtype, 5, (uint8_t*)p.dy + (size_t)p.out_coff * 2, dims, strides, box, WS_BN * 2))
      return MTB200_ERR_CUDA;
  }
  q.dw = p.dw;
  q.B = p.B; q.Do = p.Do; q.Ho = p.Ho; q.Wo = p.Wo; q.Hi = p.Hi;
  q.Cin = p.Cin; q.Cout = p.Cout;
  q.is_f16 = p.dtype == MTB200_F16;
  q.ntw = (p.Wo + q.wt - 1) / q.wt;
  q.npy = npy; q.npz = npz;
  const int npairs = npy * npz;
  const int slots_max = max(1, num_sms() / npairs);
  {
    const long long base = (long long)p.B * p.Do * q.ntw;
    double best = -1;
    int best_nhr = 1;
    for (int nhr = 1; nhr <= max(1, p.Ho / 8); ++nhr) {
      const int hlen = (p.Ho + nhr - 1) / nhr;
      if ((p.Ho + hlen - 1) / hlen != nhr) continue;
      const long long units = base * nhr;
      const long long g = units < slots_max ? units : slots_max;
      const long long per = (units + g - 1) / g;
      // every range re-reads one boundary input line
      const double eff = (double)units / (double)(per * g) * (2.0 * hlen) / (2.0 * hlen + 1.0);
      if (eff > best + 1e-9) { best = eff; best_nhr = nhr; }
    }
    q.nhr = best_nhr;
    q.hlen = (p.Ho + q.nhr - 1) / q.nhr;
  }
  q.units = (long long)p.B * p.Do * q.ntw * q.nhr;
  q.nslots = (int)(q.units < slots_max ? q.units : slots_max);
  const int smem = max(116 * 1024, q.stages * q.stage_bytes + 1024);
  cudaError_t e = cudaFuncSetAttribute(wgrad_line_s2_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("wgrad_line_s2: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  launch_pdl(wgrad_line_s2_umma_kernel, dim3(dim3((unsigned)(q.nslots * npairs))), dim3(WS_THREADS), (size_t)(smem), s, q);
  return check_launch("wgrad_line_s2_umma");
}

}  // namespace mtb
