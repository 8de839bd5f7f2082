// Fused backward of one 1x1x1 segmentation head under the MultiTalent loss (generic_UNet.py:349-351 +
// MultiTalent_Trainer_DDP.py:567-606, backward): loss pass 2 (d loss / d logits), the head's data gradient and the head's
// weight gradient in ONE pass over the voxels.
//
// A sample of a partially labelled dataset supervises 1..13 CONTIGUOUS output channels of the 47 (Task100 tables), so
// d(logits) of sample b is non-zero only inside a 16-channel window [c0_b, c0_b + 16).  The unfused sequence wrote the
// dense [voxels][48] gradient (1.5 GB at full resolution), and the pointwise data-gradient kernel and the weight-gradient
// kernel each read it back.  Here the window never leaves the SM:
//   compute warps (one voxel per thread): read the 16 logits of the window + the label, evaluate
//       d = g * ( c0 (sigma - y) - sigma (1 - sigma) (y c1 - c2) )      (mt_loss_bwd's formula, same rounding to 16 bit)
//     and write the row into a swizzled shared-memory tile  DL[128 voxels][16 channels]  (32-byte rows, SWIZZLE_32B);
//   MMA 1:  dX[128 vox][Cin]  = DL (K-major A, K = 16)  x  Wwin_b[Cin][16] (K-major B, resident)      -> TMEM, drained by
//     the same warps (transposed through shared memory so that every store instruction writes whole 128-byte lines);
//   MMA 2:  dW[Cin][16]      += X^T (MN-major A: the head's input tile, TMA)  x  DL (MN-major B: the SAME tile), K = 128
//     voxels; the accumulator stays in TMEM for the whole CTA and goes to dW[c0_b + c][ci] with fp32 atomics at the end.
// Every CTA works on tiles of ONE sample (grid = B x CTAs per sample), so window, coefficients and weight slice are
// CTA constants.  Traffic per voxel: 32..64 B of logits, 4 B label, 2 Cin B input, 2 Cin B output -- 2.8 GB instead of
// 8 GB at full resolution.
//
// Warp roles (6 warps): 0 = TMA producer (input tiles), 1 = TMEM owner + MMA issuer, 2..5 = compute + epilogue.
#include "umma.cuh"

namespace mtb {

using namespace um;

constexpr int HB_THREADS = 192;
constexpr int HB_STAGES = 4;
constexpr int HF_STAGES = 8;  // forward statistics kernel: its only HBM stream is the input tile -- keep 8 x 2 CTAs in flight per SM
constexpr int HB_WIN = 16;
constexpr int HB_MAX_LABELS = 64;

struct HeadBwdParams {
  CUtensorMap x_map;
  const void* logits;
  const float* target;
  const float4* coef;
  const float* gscale;
  const uint64_t* pos_mask;
  const void* w_swap;  // [Cin][Cout_p] 16 bit
  const void* w_fwd;   // [Cout_p][Cin] 16 bit: recompute mode (logits == nullptr) and the forward statistics kernel
  double* stats;       // forward kernel: [B][C8][4] += {sum bce, sum sigma*y, sum sigma, sum y}
  double* hard;        // forward kernel (optional): [B][C8][2] += {sum [z > 0] * y, sum [z > 0]}
  const uint64_t* valid_mask;  // forward kernel: [B] supervised channels
  void* dx;
  float* dw;           // [Cout_p][Cin] fp32
  long long nvox, tiles_per_b;
  int z_ldc, C8, n_labels, Cout_p, dx_ldc, dx_coff, accumulate, B, cps, is_f16;
  int win_c0[MTB200_MAX_HEAD_BATCH];
};

__device__ __forceinline__ void hb_tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void hb_sigmoid(float z, float& sig) {
  const float e = __expf(-fabsf(z));
  const float r = __frcp_rn(1.f + e);
  sig = z >= 0.f ? r : e * r;
}
template <typename T> __device__ __forceinline__ uint32_t hb_pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t hb_pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t hb_pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <typename T> __device__ __forceinline__ float2 hb_unpack2(uint32_t w);
template <> __device__ __forceinline__ float2 hb_unpack2<__nv_bfloat16>(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
}
template <> __device__ __forceinline__ float2 hb_unpack2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<__half2*>(&w));
}
__device__ __forceinline__ uint64_t hb_desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | (uint64_t)lo; }

// CIN = padded input channels of the head (32: full resolution, 64: second level)
// ACC: d(input) += (the level's transposed convolution already wrote its share)
// RC:  the logits were never stored (deferred head, mtb200_head_fwd_stats): MMA 0 recomputes the window
//      z[128 vox][16] = X (K-major A: the same input tile) x Wwin[16][Cin] (K-major B) one tile ahead, the compute warps read
//      it from TMEM and round it to the storage type first -- the values the stored logits would have had
template <typename T, int CIN, bool ACC, bool RC>
__global__ void __launch_bounds__(HB_THREADS, CIN == 32 ? 3 : 2) head_bwd_fused_kernel(const __grid_constant__ HeadBwdParams p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  constexpr int ROWB = CIN * 2;            // bytes per input row
  constexpr int ACT_BYTES = 128 * ROWB;    // one staged input tile
  constexpr int NCH = ROWB / 16;           // 16-byte pieces per output row
  constexpr int RPI = 32 / NCH;            // rows one warp-wide store instruction covers
  constexpr uint32_t TMEM_COLS = CIN == 32 ? 128u : 256u;  // 2 CIN (dX x 2) + 16 (dW) + 32 (RC: logits x 2)
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t act_full[HB_STAGES], act_empty[HB_STAGES];
  __shared__ __align__(8) uint64_t a_full[2], a_empty[2], d1_full[2], d1_empty[2], d2_full, z_full[2], z_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t s_pos[HB_MAX_LABELS];
  __shared__ float4 s_cf[HB_WIN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act_base = dsmem;                                     // HB_STAGES x ACT_BYTES (+ 1 KB slack: shifted M blocks)
  uint8_t* a_base = act_base + HB_STAGES * ACT_BYTES + 1024;     // 2 x 4 KB  DL tiles
  uint8_t* w_base = a_base + 2 * 4096;                           // CIN x 32 B weight window (K-major, SWIZZLE_32B)
  uint8_t* o_base = w_base + ((CIN * 32 + 1023) / 1024) * 1024;  // 4 warps x 32 rows x ROWB output transposition
  uint8_t* w0_base = o_base + 4 * 32 * ROWB;                     // RC: 16 x ROWB forward weight window (K-major, swizzled)
  constexpr uint32_t ZCOL = 2 * CIN + HB_WIN;                    // RC: two 16-column logit accumulators

  const int b = (int)blockIdx.x / p.cps, slot = (int)blockIdx.x % p.cps;
  const int c0 = p.win_c0[b];
  const long long ntiles = p.tiles_per_b;
  const bool have_work = slot < ntiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < HB_STAGES; ++i) { mbar_init(&act_full[i], 1); mbar_init(&act_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1);
      mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], 128);
      mbar_init(&z_full[i], 1); mbar_init(&z_empty[i], 128);
    }
    mbar_init(&d2_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < HB_MAX_LABELS) s_pos[threadIdx.x] = (int)threadIdx.x < p.n_labels ? p.pos_mask[threadIdx.x] : 0ull;
  if (threadIdx.x >= 64 && threadIdx.x < 64 + HB_WIN) {
    const int j = threadIdx.x - 64;
    const float gs = p.gscale ? *p.gscale : 1.f;
    float4 c = p.coef[(long long)b * p.C8 + c0 + j];
    c.x *= gs; c.y *= gs; c.z *= gs;
    s_cf[j] = c;
  }
  // weight window W[ci][c0 .. c0 + 16) -> K-major rows of 32 bytes, 16-byte pieces swizzled as TMA's SWIZZLE_32B would
  for (int idx = threadIdx.x; idx < CIN * 2; idx += HB_THREADS) {
    const int row = idx >> 1, piece = idx & 1;
    const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.w_swap) +
                                                    (long long)row * p.Cout_p + c0 + piece * 8);
    *reinterpret_cast<uint4*>(w_base + row * 32 + ((piece ^ ((row >> 2) & 1)) * 16)) = v;
  }
  if (RC) {  // forward weight rows c0 .. c0 + 16: [16][Cin], 16-byte pieces swizzled as SWIZZLE_64B / SWIZZLE_128B
    for (int idx = threadIdx.x; idx < HB_WIN * NCH; idx += HB_THREADS) {
      const int row = idx / NCH, piece = idx % NCH;
      const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.w_fwd) +
                                                      (long long)(c0 + row) * CIN + piece * 8);
      const int sw = NCH == 4 ? (piece ^ ((row >> 1) & 3)) : (piece ^ (row & 7));
      *reinterpret_cast<uint4*>(w0_base + row * ROWB + sw * 16) = v;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc(&tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: the head's input tile [128 voxels][CIN] per step =====
    uint32_t i = 0;
    for (long long t = slot; t < ntiles; t += p.cps, ++i) {
      const uint32_t stage = i % HB_STAGES;
      mbar_wait(&act_empty[stage], ((i / HB_STAGES) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&act_full[stage], (uint32_t)ACT_BYTES);
        hb_tma_load_2d(act_base + (size_t)stage * ACT_BYTES, &p.x_map, &act_full[stage], 0,
                       (int)((long long)b * p.nvox + t * 128));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t fmt = p.is_f16 ? 0u : 1u;
    const uint32_t idesc1 = idesc_f16(p.is_f16 != 0, (uint32_t)CIN, false, false);
    const uint32_t idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((HB_WIN >> 3) << 17) |
                            ((128u >> 4) << 24);
    const uint32_t act16 = __shfl_sync(0xffffffffu, (smem_u32(act_base) & 0x3FFFFu) >> 4, 0);
    const uint32_t a16 = __shfl_sync(0xffffffffu, (smem_u32(a_base) & 0x3FFFFu) >> 4, 0);
    const uint64_t w_desc = kmajor_desc(__shfl_sync(0xffffffffu, smem_u32(w_base), 0), 32u, 256u);
    // MN-major descriptors (wgrad_line.cu): SBO = 8 rows, LBO = stride between M / N blocks
    const uint32_t hi_x = ((8u * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : 4u) << 29);
    const uint32_t lbo_x = ((uint32_t)ROWB >> 4) << 16;  // M blocks past the first: row-shifted junk, never read back
    const uint32_t hi_d = ((8u * 32u) >> 4) | (1u << 14) | (6u << 29);
    const uint32_t lbo_d = (256u >> 4) << 16;
    // RC: K-major descriptors of MMA 0 (conv_pw.cu: SBO = 8 rows, 32 bytes per K step inside the swizzled row)
    const uint32_t idesc0 = idesc_f16(p.is_f16 != 0, (uint32_t)HB_WIN, false, false);
    const uint32_t hi_k = ((8u * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : 4u) << 29);
    const uint32_t w0_16 = __shfl_sync(0xffffffffu, (smem_u32(w0_base) & 0x3FFFFu) >> 4, 0);
    auto mma0 = [&](uint32_t i) {  // logits window of tile i (warp-uniform; waits for its input tile and a free z buffer)
      const uint32_t buf = i & 1u, stage = i % HB_STAGES;
      mbar_wait(&act_full[stage], (i / HB_STAGES) & 1u);
      mbar_wait(&z_empty[buf], ((i >> 1) & 1u) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t x_t = act16 + stage * ((uint32_t)ACT_BYTES >> 4);
#pragma unroll
        for (int ks = 0; ks < CIN / 16; ++ks)
          umma_f16(tmem_u + ZCOL + buf * HB_WIN, hb_desc64(hi_k, x_t + 2u * ks), hb_desc64(hi_k, w0_16 + 2u * ks), idesc0,
                   ks ? 1u : 0u);
        umma_commit(&z_full[buf]);
      }
      __syncwarp();
    };
    if (RC && have_work) mma0(0);
    uint32_t i = 0;
    for (long long t = slot; t < ntiles; t += p.cps, ++i) {
      const uint32_t buf = i & 1u, stage = i % HB_STAGES;
      if (RC && t + p.cps < ntiles) mma0(i + 1);
      mbar_wait(&a_full[buf], (i >> 1) & 1u);
      if (!RC) mbar_wait(&act_full[stage], (i / HB_STAGES) & 1u);
      mbar_wait(&d1_empty[buf], ((i >> 1) & 1u) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_t = a16 + buf * (4096u >> 4);
        umma_f16(tmem_u + buf * CIN, kmajor_desc((a_t << 4), 32u, 256u), w_desc, idesc1, 0u);
        umma_commit(&d1_full[buf]);
        const uint32_t x_t = act16 + stage * ((uint32_t)ACT_BYTES >> 4);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16(tmem_u + 2 * CIN, hb_desc64(hi_x, (x_t + (uint32_t)(kk * ROWB)) | lbo_x),
                   hb_desc64(hi_d, (a_t + (uint32_t)(kk * 32)) | lbo_d), idesc2, (i > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&act_empty[stage]);
        umma_commit(&a_empty[buf]);
      }
      __syncwarp();
    }
    if (have_work) {
      if (elect_one()) umma_commit(&d2_full);
      __syncwarp();
    }
  } else if (have_work) {
    // ===== compute + epilogue warps: TMEM lane m = voxel row m of the tile =====
    const int q = warp & 3;
    const int m = q * 32 + lane;
    unsigned vbits = 0;
#pragma unroll
    for (int j = 0; j < HB_WIN; ++j)
      if (s_cf[j].w != 0.f) vbits |= 1u << j;
    const T* zb = reinterpret_cast<const T*>(p.logits) + (long long)b * p.nvox * p.z_ldc + c0;
    const float* tb = p.target + (long long)b * p.nvox;
    T* dxb = reinterpret_cast<T*>(p.dx) + (long long)b * p.nvox * p.dx_ldc + p.dx_coff;
    uint8_t* o_warp = o_base + q * (32 * ROWB);

    Raw8<T> z0, z1, n0, n1;
    float lab = 0.f, nlab = 0.f;
    auto fetch = [&](long long t, Raw8<T>& a, Raw8<T>& c, float& l) {
      const long long v = min(t * 128 + m, p.nvox - 1);  // rows past the sample: clamped re-read, masked below
      if (!RC) {
        a.load(zb + v * p.z_ldc);
        c.load(zb + v * p.z_ldc + 8);
      }
      l = __ldg(tb + v);
    };
    auto epilogue = [&](uint32_t i, long long t) {
      const uint32_t buf = i & 1u;
      // accumulate: the other consumer's share of d(input), requested before the wait for the tensor core so that its
      // DRAM latency overlaps it (same rows / pieces as the write-out below)
      uint4 old[ACC ? NCH : 1];
      if (ACC) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const long long v = min(t * 128 + q * 32 + j * RPI + lane / NCH, p.nvox - 1);
          const T* src = dxb + v * p.dx_ldc + (lane % NCH) * 8;
          asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(old[j].x), "=r"(old[j].y), "=r"(old[j].z), "=r"(old[j].w) : "l"(src));
        }
      }
      mbar_wait(&d1_full[buf], (i >> 1) & 1u);
      tc_fence_after();
      // accumulator row -> 16-bit -> this warp's staging rows (pieces XOR-swizzled: conflict-free both ways)
#pragma unroll
      for (int c16 = 0; c16 < CIN / 16; ++c16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * CIN + (uint32_t)(c16 * 16), r);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint4 w;
          w.x = hb_pack2<T>(__uint_as_float(r[8 * h + 0]), __uint_as_float(r[8 * h + 1]));
          w.y = hb_pack2<T>(__uint_as_float(r[8 * h + 2]), __uint_as_float(r[8 * h + 3]));
          w.z = hb_pack2<T>(__uint_as_float(r[8 * h + 4]), __uint_as_float(r[8 * h + 5]));
          w.w = hb_pack2<T>(__uint_as_float(r[8 * h + 6]), __uint_as_float(r[8 * h + 7]));
          const int piece = c16 * 2 + h;
          const int sw = NCH == 4 ? (piece ^ ((lane >> 1) & 3)) : (piece ^ (lane & 7));
          *reinterpret_cast<uint4*>(o_warp + lane * ROWB + sw * 16) = w;
        }
      }
      tc_fence_before();
      mbar_arrive(&d1_empty[buf]);
      __syncwarp();
      // write-out: every instruction covers RPI whole rows = 512 contiguous bytes (when dx_ldc == CIN)
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const int row = j * RPI + lane / NCH, piece = lane % NCH;
        const int sw = NCH == 4 ? (piece ^ ((row >> 1) & 3)) : (piece ^ (row & 7));
        uint4 w = *reinterpret_cast<const uint4*>(o_warp + row * ROWB + sw * 16);
        const long long v = t * 128 + q * 32 + row;
        if (v < p.nvox) {
          uint4* dst = reinterpret_cast<uint4*>(dxb + v * p.dx_ldc + piece * 8);
          if (ACC) {
            const uint32_t* ow = reinterpret_cast<const uint32_t*>(&old[ACC ? j : 0]);
            uint32_t* ww = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 a = hb_unpack2<T>(ow[e]), c = hb_unpack2<T>(ww[e]);
              ww[e] = hb_pack2<T>(a.x + c.x, a.y + c.y);
            }
          }
          *dst = w;
        }
      }
      __syncwarp();
    };

    fetch(slot, z0, z1, lab);
    uint32_t i = 0;
    long long tprev = 0;
    for (long long t = slot; t < ntiles; t += p.cps, ++i) {
      const uint32_t buf = i & 1u;
      // ---- d(logits) row of this voxel
      const bool inside = t * 128 + m < p.nvox;
      const int li = (int)lab;
      const uint64_t pm = (inside && (unsigned)li < (unsigned)HB_MAX_LABELS) ? s_pos[li] : 0ull;
      const unsigned ybits = (unsigned)((pm >> c0) & 0xffffull);
      uint32_t zr[16];
      if (RC) {
        mbar_wait(&z_full[buf], (i >> 1) & 1u);
        tc_fence_after();
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + ZCOL + buf * HB_WIN, zr);
        tc_fence_before();
        mbar_arrive(&z_empty[buf]);
      }
      float d[HB_WIN];
#pragma unroll
      for (int j = 0; j < HB_WIN; ++j) {
        d[j] = 0.f;
        if (vbits & (1u << j)) {
          const float zz = RC ? Traits<T>::round(__uint_as_float(zr[j])) : (j < 8 ? z0.get(j) : z1.get(j - 8));
          const float y = (ybits >> j) & 1u ? 1.f : 0.f;
          const float4 cf = s_cf[j];
          float sig;
          hb_sigmoid(zz, sig);
          d[j] = inside ? cf.x * (sig - y) - sig * (1.f - sig) * (y * cf.y - cf.z) : 0.f;
        }
      }
      uint4 w0, w1;
      w0.x = hb_pack2<T>(d[0], d[1]); w0.y = hb_pack2<T>(d[2], d[3]); w0.z = hb_pack2<T>(d[4], d[5]); w0.w = hb_pack2<T>(d[6], d[7]);
      w1.x = hb_pack2<T>(d[8], d[9]); w1.y = hb_pack2<T>(d[10], d[11]); w1.z = hb_pack2<T>(d[12], d[13]); w1.w = hb_pack2<T>(d[14], d[15]);
      mbar_wait(&a_empty[buf], ((i >> 1) & 1u) ^ 1u);
      uint8_t* arow = a_base + buf * 4096 + m * 32;
      const int s = (m >> 2) & 1;
      *reinterpret_cast<uint4*>(arow + (s ? 16 : 0)) = w0;
      *reinterpret_cast<uint4*>(arow + (s ? 0 : 16)) = w1;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&a_full[buf]);
      // ---- next tile's loads in flight under the previous tile's epilogue
      const long long tn = t + p.cps;
      if (tn < ntiles) fetch(tn, n0, n1, nlab);
      if (i > 0) epilogue(i - 1, tprev);
      tprev = t;
      z0 = n0; z1 = n1; lab = nlab;
    }
    epilogue(i - 1, tprev);
    // ---- weight gradient of this CTA: D2[ci][c] -> dW[c0 + c][ci]
    if (q * 32 < CIN) {
      mbar_wait(&d2_full, 0);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + 2 * CIN, r);
      float* dst = p.dw + (long long)c0 * CIN + m;
#pragma unroll
      for (int e = 0; e < HB_WIN; ++e) {
        const float v = __uint_as_float(r[e]);
        if (v != 0.f) atomicAdd(dst + (long long)e * CIN, v);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- forward of a deferred head: loss pass 1 straight from the head's input ----------------------------------------------
// z[128 vox][16] = X x Wwin^T on the tensor core (the window of the sample's dataset), rounded to the storage type, then
// mt_loss_stats' arithmetic per supervised channel; the logits are never written.  Per-thread partial sums live in
// registers for the whole CTA (one sample per CTA), are combined in a fixed order inside the CTA and leave as one fp64
// atomic per (channel, statistic).  HARD: also the thresholded-prediction counts of run_online_evaluation.
template <typename T, int CIN, bool HARD>
__global__ void __launch_bounds__(HB_THREADS, 2) head_fwd_stats_kernel(const __grid_constant__ HeadBwdParams p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  constexpr int ROWB = CIN * 2;
  constexpr int ACT_BYTES = 128 * ROWB;
  constexpr int NCH = ROWB / 16;
  constexpr int NQ = HARD ? 6 : 4;
  constexpr int NST = CIN == 32 ? HF_STAGES : 6;  // 64-channel inputs: 6 x 16 KB so that two CTAs still share an SM
  extern __shared__ uint8_t dsmem_raw[];
  constexpr int NZ = 8;  // logit accumulators in flight (16 TMEM columns each): the MMA -> TMEM -> threads -> MMA round trip
                         // is ~1 us, two buffers kept the tensor core idle most of the time
  __shared__ __align__(8) uint64_t act_full[HF_STAGES], act_empty[HF_STAGES], z_full[NZ], z_empty[NZ];
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t s_pos[HB_MAX_LABELS];
  __shared__ float s_part[4][NQ * HB_WIN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act_base = dsmem;
  uint8_t* w0_base = act_base + NST * ACT_BYTES;
  const int b = (int)blockIdx.x / p.cps, slot = (int)blockIdx.x % p.cps;
  const int c0 = p.win_c0[b];
  const long long ntiles = p.tiles_per_b;
  const bool have_work = slot < ntiles;
  const unsigned vbits = (unsigned)((p.valid_mask[b] >> c0) & 0xffffull);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(&act_full[i], 1); mbar_init(&act_empty[i], 1); }
    for (int i = 0; i < NZ; ++i) { mbar_init(&z_full[i], 1); mbar_init(&z_empty[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < HB_MAX_LABELS) s_pos[threadIdx.x] = (int)threadIdx.x < p.n_labels ? p.pos_mask[threadIdx.x] : 0ull;
  for (int idx = threadIdx.x; idx < HB_WIN * NCH; idx += HB_THREADS) {
    const int row = idx / NCH, piece = idx % NCH;
    const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.w_fwd) +
                                                    (long long)(c0 + row) * CIN + piece * 8);
    const int sw = NCH == 4 ? (piece ^ ((row >> 1) & 3)) : (piece ^ (row & 7));
    *reinterpret_cast<uint4*>(w0_base + row * ROWB + sw * 16) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc(&tmem_slot, (uint32_t)(NZ * HB_WIN));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    uint32_t i = 0;
    for (long long t = slot; t < ntiles; t += p.cps, ++i) {
      const uint32_t stage = i % NST;
      mbar_wait(&act_empty[stage], ((i / NST) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&act_full[stage], (uint32_t)ACT_BYTES);
        hb_tma_load_2d(act_base + (size_t)stage * ACT_BYTES, &p.x_map, &act_full[stage], 0,
                       (int)((long long)b * p.nvox + t * 128));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc0 = idesc_f16(p.is_f16 != 0, (uint32_t)HB_WIN, false, false);
    const uint32_t hi_k = ((8u * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : 4u) << 29);
    const uint32_t act16 = __shfl_sync(0xffffffffu, (smem_u32(act_base) & 0x3FFFFu) >> 4, 0);
    const uint32_t w0_16 = __shfl_sync(0xffffffffu, (smem_u32(w0_base) & 0x3FFFFu) >> 4, 0);
    uint32_t i = 0;
    for (long long t = slot; t < ntiles; t += p.cps, ++i) {
      const uint32_t buf = i % NZ, stage = i % NST;
      mbar_wait(&act_full[stage], (i / NST) & 1u);
      mbar_wait(&z_empty[buf], ((i / NZ) & 1u) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t x_t = act16 + stage * ((uint32_t)ACT_BYTES >> 4);
#pragma unroll
        for (int ks = 0; ks < CIN / 16; ++ks)
          umma_f16(tmem_u + buf * HB_WIN, hb_desc64(hi_k, x_t + 2u * ks), hb_desc64(hi_k, w0_16 + 2u * ks), idesc0,
                   ks ? 1u : 0u);
        umma_commit(&z_full[buf]);
        umma_commit(&act_empty[stage]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    float part[NQ][HB_WIN];
#pragma unroll
    for (int a = 0; a < NQ; ++a)
#pragma unroll
      for (int j = 0; j < HB_WIN; ++j) part[a][j] = 0.f;
    const float* tb = p.target + (long long)b * p.nvox;
    if (have_work) {
      float lab = __ldg(tb + min((long long)slot * 128 + m, p.nvox - 1));
      uint32_t i = 0;
      for (long long t = slot; t < ntiles; t += p.cps, ++i) {
        const uint32_t buf = i % NZ;
        const long long tn = t + p.cps;
        float nlab = 0.f;
        if (tn < ntiles) nlab = __ldg(tb + min(tn * 128 + m, p.nvox - 1));
        uint32_t zr[16];
        mbar_wait(&z_full[buf], (i / NZ) & 1u);
        tc_fence_after();
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + buf * HB_WIN, zr);
        tc_fence_before();
        mbar_arrive(&z_empty[buf]);
        if (t * 128 + m < p.nvox) {
          const int li = (int)lab;
          const uint64_t pm = (unsigned)li < (unsigned)HB_MAX_LABELS ? s_pos[li] : 0ull;
          const unsigned ybits = (unsigned)((pm >> c0) & 0xffffull);
#pragma unroll
          for (int j = 0; j < HB_WIN; ++j) {
            if (vbits & (1u << j)) {
              const float zz = Traits<T>::round(__uint_as_float(zr[j]));
              const float y = (ybits >> j) & 1u ? 1.f : 0.f;
              const float e = __expf(-fabsf(zz));
              const float r = __frcp_rn(1.f + e);
              const float sig = zz >= 0.f ? r : e * r;
              const float l1p = e < 2.44140625e-4f ? e * (1.f - 0.5f * e) : __logf(1.f + e);
              part[0][j] += fmaxf(zz, 0.f) - zz * y + l1p;
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
        lab = nlab;
      }
    }
    // lanes -> warp -> CTA in a fixed order, then one double atomic per (channel, statistic)
#pragma unroll
    for (int a = 0; a < NQ; ++a)
#pragma unroll
      for (int j = 0; j < HB_WIN; ++j) {
        const float v = warp_sum(part[a][j]);
        if (lane == 0) s_part[q][a * HB_WIN + j] = v;
      }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int idx = (warp - 2) * 32 + lane;
    if (idx < NQ * HB_WIN) {
      const int a = idx / HB_WIN, j = idx % HB_WIN;
      const float v = s_part[0][idx] + s_part[1][idx] + s_part[2][idx] + s_part[3][idx];
      if ((vbits & (1u << j)) && v != 0.f) {
        if (a < 4) atomicAdd(p.stats + ((long long)b * p.C8 + c0 + j) * 4 + a, (double)v);
        else atomicAdd(p.hard + ((long long)b * p.C8 + c0 + j) * 2 + (a - 4), (double)v);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)(NZ * HB_WIN));
  }
}

// ---- inference: head -> non-linearity x Gaussian weight -> scatter-add into the sliding-window accumulators ------------------
// (generic_UNet.py:349-351 + neural_network.py:374-394, 531-589.)  The logits of a tile never reach memory: MMA (N = padded
// class count) -> TMEM -> one thread per SOURCE voxel: + bias, round to the storage type (the values the stored logits had),
// sigmoid / softmax / identity, x weight x gauss[destination voxel], read-modify-write of the C channel-first accumulators at
// the un-flipped destination (a warp covers 32 consecutive w voxels: 128-byte rows per class).  One launch per tile, tiles
// in stream order (overlapping tiles must not race; the reference accumulates in tile order too).
constexpr int HA_STAGES = 6;
struct HeadAggParams {
  CUtensorMap x_map;
  const void* w_fwd;      // [Cout_p][Cin]
  const float* bias;      // [Cout_p] or null
  const float* gauss;     // [pd][ph][pw] or null
  float* acc;             // [C][X][Y][Z]
  float* nb;              // [X][Y][Z] or null
  float weight;
  int C, Cout_p, pd, ph, pw, flip, nonlin, X, Y, Z, x0, y0, z0, ntiles, is_f16;
  long long nvox;
};

template <typename T, int CIN>
__global__ void __launch_bounds__(HB_THREADS, 2) head_aggregate_kernel(const __grid_constant__ HeadAggParams p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  constexpr int ROWB = CIN * 2;
  constexpr int ACT_BYTES = 128 * ROWB;
  constexpr int NCH = ROWB / 16;
  constexpr int NP = 48;  // padded class count (TMEM columns per accumulator)
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t act_full[HA_STAGES], act_empty[HA_STAGES], z_full[2], z_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[NP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act_base = dsmem;
  uint8_t* w_base = act_base + HA_STAGES * ACT_BYTES;

  if (threadIdx.x == 0) {
    for (int i = 0; i < HA_STAGES; ++i) { mbar_init(&act_full[i], 1); mbar_init(&act_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&z_full[i], 1); mbar_init(&z_empty[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < NP) s_bias[threadIdx.x] = (p.bias && (int)threadIdx.x < p.Cout_p) ? p.bias[threadIdx.x] : 0.f;
  for (int idx = threadIdx.x; idx < NP * NCH; idx += HB_THREADS) {
    const int row = idx / NCH, piece = idx % NCH;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row < p.Cout_p)
      v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.w_fwd) + (long long)row * CIN + piece * 8);
    const int sw = NCH == 4 ? (piece ^ ((row >> 1) & 3)) : (piece ^ (row & 7));
    *reinterpret_cast<uint4*>(w_base + row * ROWB + sw * 16) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc(&tmem_slot, 128u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    uint32_t i = 0;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++i) {
      const uint32_t stage = i % HA_STAGES;
      mbar_wait(&act_empty[stage], ((i / HA_STAGES) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&act_full[stage], (uint32_t)ACT_BYTES);
        hb_tma_load_2d(act_base + (size_t)stage * ACT_BYTES, &p.x_map, &act_full[stage], 0, t * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc0 = idesc_f16(p.is_f16 != 0, (uint32_t)NP, false, false);
    const uint32_t hi_k = ((8u * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : 4u) << 29);
    const uint32_t act16 = __shfl_sync(0xffffffffu, (smem_u32(act_base) & 0x3FFFFu) >> 4, 0);
    const uint32_t w16 = __shfl_sync(0xffffffffu, (smem_u32(w_base) & 0x3FFFFu) >> 4, 0);
    uint32_t i = 0;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++i) {
      const uint32_t buf = i & 1u, stage = i % HA_STAGES;
      mbar_wait(&act_full[stage], (i / HA_STAGES) & 1u);
      mbar_wait(&z_empty[buf], ((i >> 1) & 1u) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t x_t = act16 + stage * ((uint32_t)ACT_BYTES >> 4);
#pragma unroll
        for (int ks = 0; ks < CIN / 16; ++ks)
          umma_f16(tmem_u + buf * 64u, hb_desc64(hi_k, x_t + 2u * ks), hb_desc64(hi_k, w16 + 2u * ks), idesc0, ks ? 1u : 0u);
        umma_commit(&z_full[buf]);
        umma_commit(&act_empty[stage]);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const long long plane = (long long)p.Y * p.Z, cstride = (long long)p.X * plane;
    uint32_t i = 0;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++i) {
      const uint32_t buf = i & 1u;
      // source voxel of this thread and its un-flipped destination
      const uint32_t v = (uint32_t)t * 128u + (uint32_t)m;  // nvox < 2^31 (host check): 32-bit divisions
      const bool inside = (long long)v < p.nvox;
      const uint32_t rem = v / (uint32_t)p.pw;
      const int sw_ = (int)(v - rem * (uint32_t)p.pw);
      const int sd_ = (int)(rem / (uint32_t)p.ph), sh_ = (int)(rem - (uint32_t)sd_ * (uint32_t)p.ph);
      const int d = (p.flip & 4) ? p.pd - 1 - sd_ : sd_, h = (p.flip & 2) ? p.ph - 1 - sh_ : sh_;
      const int w = (p.flip & 1) ? p.pw - 1 - sw_ : sw_;
      float gw = p.weight;
      if (inside && p.gauss) gw *= __ldg(p.gauss + ((long long)d * p.ph + h) * p.pw + w);
      uint32_t r[3][16];
      mbar_wait(&z_full[buf], (i >> 1) & 1u);
      tc_fence_after();
      const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 64u;
      tmem_ld16_async(tl, r[0]);
      tmem_ld16_async(tl + 16u, r[1]);
      tmem_ld16_async(tl + 32u, r[2]);
      tmem_ld_fence(r[0]); tmem_ld_fence(r[1]); tmem_ld_fence(r[2]);
      tc_fence_before();
      mbar_arrive(&z_empty[buf]);
      if (!inside) continue;
      float val[NP];
#pragma unroll
      for (int c = 0; c < NP; ++c) val[c] = Traits<T>::round(__uint_as_float(r[c >> 4][c & 15]) + s_bias[c]);
      if (p.nonlin == 1) {
#pragma unroll
        for (int c = 0; c < NP; ++c) {
          const float e = expf(-fabsf(val[c]));
          val[c] = (val[c] >= 0.f ? 1.f / (1.f + e) : e / (1.f + e)) * gw;
        }
      } else if (p.nonlin == 2) {
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < NP; ++c) if (c < p.C) mx = fmaxf(mx, val[c]);
        float ssum = 0.f;
#pragma unroll
        for (int c = 0; c < NP; ++c) { val[c] = c < p.C ? expf(val[c] - mx) : 0.f; ssum += val[c]; }
        const float g2 = gw / ssum;
#pragma unroll
        for (int c = 0; c < NP; ++c) val[c] *= g2;
      } else {
#pragma unroll
        for (int c = 0; c < NP; ++c) val[c] *= gw;
      }
      float* dst = p.acc + ((long long)(p.x0 + d) * p.Y + (p.y0 + h)) * p.Z + p.z0 + w;
      // fire-and-forget reductions (RED.ADD.F32): the tiles of one launch never overlap and launches are stream ordered, so
      // every location receives exactly one add per launch -- the same sums as a read-modify-write, without holding a
      // register per outstanding load (47 classes x 128 threads in flight per CTA)
#pragma unroll
      for (int c = 0; c < NP; ++c)
        if (c < p.C) atomicAdd(dst + (long long)c * cstride, val[c]);
      if (p.nb) {
        float* nbp = p.nb + ((long long)(p.x0 + d) * p.Y + (p.y0 + h)) * p.Z + p.z0 + w;
        atomicAdd(nbp, p.gauss ? __ldg(p.gauss + ((long long)d * p.ph + h) * p.pw + w) : 1.f);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128u);
  }
}

int umma_available();

int head_bwd_fused(const mtb200_head_bwd_params& p, cudaStream_t s) {
  if (!umma_available()) { set_error("head_bwd_fused: no sm_100 device / driver entry point"); return MTB200_ERR_UNSUPPORTED; }
  if (p.dtype != MTB200_BF16 && p.dtype != MTB200_F16) { set_error("head_bwd_fused: 16-bit tensors only"); return MTB200_ERR_UNSUPPORTED; }
  if (p.Cin != 32 && p.Cin != 64) { set_error("head_bwd_fused: Cin %d (32 or 64)", p.Cin); return MTB200_ERR_UNSUPPORTED; }
  MTB_REQUIRE(p.B >= 1 && p.B <= MTB200_MAX_HEAD_BATCH, "head_bwd_fused: batch %d (max %d)", p.B, MTB200_MAX_HEAD_BATCH);
  MTB_REQUIRE(p.z_ldc % 8 == 0 && p.x_ldc % 8 == 0 && p.x_coff % 8 == 0 && p.dx_ldc % 8 == 0 && p.dx_coff % 8 == 0 &&
                  p.Cout % 8 == 0 && p.C8 % 8 == 0,
              "head_bwd_fused: strides / offsets must be multiples of 8 channels");
  MTB_REQUIRE(p.n_labels <= HB_MAX_LABELS, "head_bwd_fused: n_labels=%d > %d", p.n_labels, HB_MAX_LABELS);
  MTB_REQUIRE(p.nvox > 0 && (long long)p.B * p.nvox < (1LL << 31), "head_bwd_fused: %lld voxels", (long long)p.nvox);
  for (int b = 0; b < p.B; ++b)
    MTB_REQUIRE(p.win_c0[b] >= 0 && p.win_c0[b] % 8 == 0 && p.win_c0[b] + HB_WIN <= p.Cout && p.win_c0[b] + HB_WIN <= p.C8 &&
                    p.win_c0[b] + HB_WIN <= p.z_ldc,
                "head_bwd_fused: window of sample %d starts at channel %d", b, p.win_c0[b]);
  static thread_local HeadBwdParams q;
  memset(&q, 0, sizeof(q));
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)((long long)p.B * p.nvox)};
    cuuint64_t strides[1] = {(cuuint64_t)p.x_ldc * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.Cin, 128};
    if (!umma_encode_map(&q.x_map, p.dtype, 2, (uint8_t*)p.x + (size_t)p.x_coff * 2, dims, strides, box, p.Cin * 2))
      return MTB200_ERR_CUDA;
  }
  q.w_fwd = p.w_fwd;
  q.logits = p.logits; q.target = p.target; q.coef = reinterpret_cast<const float4*>(p.coef); q.gscale = p.gscale;
  q.pos_mask = p.pos_mask; q.w_swap = p.w_swap; q.dx = p.dx; q.dw = p.dw;
  q.nvox = p.nvox; q.tiles_per_b = (p.nvox + 127) / 128;
  q.z_ldc = p.z_ldc; q.C8 = p.C8; q.n_labels = p.n_labels; q.Cout_p = p.Cout; q.dx_ldc = p.dx_ldc; q.dx_coff = p.dx_coff;
  q.accumulate = p.accumulate; q.B = p.B; q.is_f16 = p.dtype == MTB200_F16;
  for (int b = 0; b < p.B; ++b) q.win_c0[b] = p.win_c0[b];
  const int per_sm = p.Cin == 32 ? 3 : 2;
  long long cps = ((long long)num_sms() * per_sm) / p.B;
  if (cps < 1) cps = 1;
  if (cps > q.tiles_per_b) cps = q.tiles_per_b;
  q.cps = (int)cps;
  const int rowb = p.Cin * 2;
  const int smem = HB_STAGES * 128 * rowb + 1024 + 2 * 4096 + ((p.Cin * 32 + 1023) / 1024) * 1024 + 4 * 32 * rowb +
                   HB_WIN * rowb + 1024;
  dim3 grid((unsigned)(q.cps * p.B));
  cudaError_t e = cudaSuccess;
#define HB_LAUNCH3(T, CIN, ACC, RC)                                                                                      \
  do {                                                                                                                   \
    e = cudaFuncSetAttribute(head_bwd_fused_kernel<T, CIN, ACC, RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    if (e == cudaSuccess) launch_pdl(head_bwd_fused_kernel<T, CIN, ACC, RC>, dim3(grid), dim3(HB_THREADS), (size_t)(smem), s, q);                      \
  } while (0)
#define HB_LAUNCH2(T, CIN, ACC)                                                   \
  do {                                                                            \
    if (p.logits) HB_LAUNCH3(T, CIN, ACC, false); else HB_LAUNCH3(T, CIN, ACC, true); \
  } while (0)
#define HB_LAUNCH(T, CIN)                                                     \
  do {                                                                        \
    if (p.accumulate) HB_LAUNCH2(T, CIN, true); else HB_LAUNCH2(T, CIN, false); \
  } while (0)
  if (p.dtype == MTB200_BF16) {
    if (p.Cin == 32) HB_LAUNCH(__nv_bfloat16, 32); else HB_LAUNCH(__nv_bfloat16, 64);
  } else {
    if (p.Cin == 32) HB_LAUNCH(__half, 32); else HB_LAUNCH(__half, 64);
  }
#undef HB_LAUNCH2
#undef HB_LAUNCH3
#undef HB_LAUNCH
  if (e != cudaSuccess) { set_error("head_bwd_fused: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  return check_launch("head_bwd_fused");
}

int head_fwd_stats(const mtb200_head_fwd_params& p, cudaStream_t s) {
  if (!umma_available()) { set_error("head_fwd_stats: no sm_100 device / driver entry point"); return MTB200_ERR_UNSUPPORTED; }
  if (p.dtype != MTB200_BF16 && p.dtype != MTB200_F16) { set_error("head_fwd_stats: 16-bit tensors only"); return MTB200_ERR_UNSUPPORTED; }
  if (p.Cin != 32 && p.Cin != 64) { set_error("head_fwd_stats: Cin %d (32 or 64)", p.Cin); return MTB200_ERR_UNSUPPORTED; }
  MTB_REQUIRE(p.B >= 1 && p.B <= MTB200_MAX_HEAD_BATCH, "head_fwd_stats: batch %d (max %d)", p.B, MTB200_MAX_HEAD_BATCH);
  MTB_REQUIRE(p.x_ldc % 8 == 0 && p.x_coff % 8 == 0 && p.Cout % 8 == 0 && p.C8 % 8 == 0, "head_fwd_stats: alignment");
  MTB_REQUIRE(p.n_labels <= HB_MAX_LABELS, "head_fwd_stats: n_labels=%d > %d", p.n_labels, HB_MAX_LABELS);
  MTB_REQUIRE(p.nvox > 0 && (long long)p.B * p.nvox < (1LL << 31), "head_fwd_stats: %lld voxels", (long long)p.nvox);
  for (int b = 0; b < p.B; ++b)
    MTB_REQUIRE(p.win_c0[b] >= 0 && p.win_c0[b] % 8 == 0 && p.win_c0[b] + HB_WIN <= p.Cout && p.win_c0[b] + HB_WIN <= p.C8,
                "head_fwd_stats: window of sample %d starts at channel %d", b, p.win_c0[b]);
  static thread_local HeadBwdParams q;
  memset(&q, 0, sizeof(q));
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)((long long)p.B * p.nvox)};
    cuuint64_t strides[1] = {(cuuint64_t)p.x_ldc * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.Cin, 128};
    if (!umma_encode_map(&q.x_map, p.dtype, 2, (uint8_t*)p.x + (size_t)p.x_coff * 2, dims, strides, box, p.Cin * 2))
      return MTB200_ERR_CUDA;
  }
  q.target = p.target; q.pos_mask = p.pos_mask; q.valid_mask = p.valid_mask; q.w_fwd = p.w_fwd; q.stats = p.stats; q.hard = p.hard;
  q.nvox = p.nvox; q.tiles_per_b = (p.nvox + 127) / 128;
  q.C8 = p.C8; q.n_labels = p.n_labels; q.Cout_p = p.Cout; q.B = p.B; q.is_f16 = p.dtype == MTB200_F16;
  for (int b = 0; b < p.B; ++b) q.win_c0[b] = p.win_c0[b];
  long long cps = ((long long)num_sms() * 2) / p.B;  // two CTAs per SM (the partial sums take ~100 registers per thread)
  if (cps < 1) cps = 1;
  if (cps > q.tiles_per_b) cps = q.tiles_per_b;
  q.cps = (int)cps;
  const int rowb = p.Cin * 2;
  const int smem = (p.Cin == 32 ? HF_STAGES : 6) * 128 * rowb + HB_WIN * rowb + 2048;
  dim3 grid((unsigned)(q.cps * p.B));
  cudaError_t e = cudaSuccess;
#define HF_LAUNCH2(T, CIN, HARD)                                                                                     \
  do {                                                                                                               \
    e = cudaFuncSetAttribute(head_fwd_stats_kernel<T, CIN, HARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    if (e == cudaSuccess) launch_pdl(head_fwd_stats_kernel<T, CIN, HARD>, dim3(grid), dim3(HB_THREADS), (size_t)(smem), s, q);                     \
  } while (0)
#define HF_LAUNCH(T, CIN)                                                \
  do {                                                                   \
    if (p.hard) HF_LAUNCH2(T, CIN, true); else HF_LAUNCH2(T, CIN, false); \
  } while (0)
  if (p.dtype == MTB200_BF16) {
    if (p.Cin == 32) HF_LAUNCH(__nv_bfloat16, 32); else HF_LAUNCH(__nv_bfloat16, 64);
  } else {
    if (p.Cin == 32) HF_LAUNCH(__half, 32); else HF_LAUNCH(__half, 64);
  }
#undef HF_LAUNCH
#undef HF_LAUNCH2
  if (e != cudaSuccess) { set_error("head_fwd_stats: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  return check_launch("head_fwd_stats");
}

int head_aggregate(const mtb200_head_agg_params& p, cudaStream_t s) {
  if (!umma_available()) { set_error("head_aggregate: no sm_100 device / driver entry point"); return MTB200_ERR_UNSUPPORTED; }
  if (p.dtype != MTB200_BF16 && p.dtype != MTB200_F16) { set_error("head_aggregate: 16-bit tensors only"); return MTB200_ERR_UNSUPPORTED; }
  if (p.Cin != 32 && p.Cin != 64) { set_error("head_aggregate: Cin %d (32 or 64)", p.Cin); return MTB200_ERR_UNSUPPORTED; }
  if (p.Cout > 48 || p.Cout % 8 || p.C > p.Cout) { set_error("head_aggregate: %d classes (padded %d; at most 48)", p.C, p.Cout); return MTB200_ERR_UNSUPPORTED; }
  MTB_REQUIRE(p.x_ldc % 8 == 0 && p.x_coff % 8 == 0, "head_aggregate: alignment");
  MTB_REQUIRE(p.x0 >= 0 && p.y0 >= 0 && p.z0 >= 0 && p.x0 + p.pd <= p.X && p.y0 + p.ph <= p.Y && p.z0 + p.pw <= p.Z,
              "head_aggregate: tile outside the volume");
  const long long nvox = (long long)p.pd * p.ph * p.pw;
  MTB_REQUIRE(nvox > 0 && nvox < (1LL << 31), "head_aggregate: %lld voxels", nvox);
  static thread_local HeadAggParams q;
  memset(&q, 0, sizeof(q));
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)nvox};
    cuuint64_t strides[1] = {(cuuint64_t)p.x_ldc * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.Cin, 128};
    if (!umma_encode_map(&q.x_map, p.dtype, 2, (uint8_t*)p.x + (size_t)p.x_coff * 2, dims, strides, box, p.Cin * 2))
      return MTB200_ERR_CUDA;
  }
  q.w_fwd = p.w_fwd; q.bias = p.bias; q.gauss = p.gauss; q.acc = p.acc; q.nb = p.nb; q.weight = p.weight;
  q.C = p.C; q.Cout_p = p.Cout; q.pd = p.pd; q.ph = p.ph; q.pw = p.pw; q.flip = p.flip; q.nonlin = p.nonlin;
  q.X = p.X; q.Y = p.Y; q.Z = p.Z; q.x0 = p.x0; q.y0 = p.y0; q.z0 = p.z0;
  q.nvox = nvox; q.ntiles = (int)((nvox + 127) / 128); q.is_f16 = p.dtype == MTB200_F16;
  const int rowb = p.Cin * 2;
  const int smem = HA_STAGES * 128 * rowb + 48 * rowb + 2048;
  const int gx = min(q.ntiles, 2 * num_sms());
  cudaError_t e = cudaSuccess;
#define HA_LAUNCH(T, CIN)                                                                                          \
  do {                                                                                                             \
    e = cudaFuncSetAttribute(head_aggregate_kernel<T, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);    \
    if (e == cudaSuccess) launch_pdl(head_aggregate_kernel<T, CIN>, dim3(gx), dim3(HB_THREADS), (size_t)(smem), s, q);                           \
  } while (0)
  if (p.dtype == MTB200_BF16) {
    if (p.Cin == 32) HA_LAUNCH(__nv_bfloat16, 32); else HA_LAUNCH(__nv_bfloat16, 64);
  } else {
    if (p.Cin == 32) HA_LAUNCH(__half, 32); else HA_LAUNCH(__half, 64);
  }
#undef HA_LAUNCH
  if (e != cudaSuccess) { set_error("head_aggregate: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  return check_launch("head_aggregate");
}

}  // namespace mtb
