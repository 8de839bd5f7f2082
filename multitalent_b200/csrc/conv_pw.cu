// Pointwise (1x1x1) tcgen05 convolution as a flat streaming GEMM -- the segmentation heads (generic_UNet.py:349-351,
// 30..320 -> 47 channels on every decoder level) and their data gradient (47 -> 30..320).  These layers do ~50 FLOP per
// byte: they are HBM streams, so the kernel is built around the memory system rather than the tensor pipe:
//   * the NDHWC tensor is a plain matrix [voxels][channels]; a tile is 128 CONSECUTIVE voxels (no 3-D bricks, no halo),
//     fetched by one 2-D TMA box per 64-channel chunk (K-major, 128/64/32-byte swizzle; channels past Cin and rows past
//     the end of the tensor are zero-filled by TMA, so Cin = 48 runs as one K = 64 chunk);
//   * the whole weight matrix stays resident in shared memory (loaded once per CTA);
//   * tcgen05.mma M = 128, N = Cout, two TMEM accumulators (tile k drains while tile k+1 accumulates);
//   * epilogue: TMEM -> registers -> (+bias, round) -> swizzled shared-memory staging tile -> ONE TMA store per column
//     block.  The per-tap kernel's direct stores (one 16-byte piece per lane at a row-pitch stride: 32 half-filled sectors
//     per instruction) were what limited it to ~3 TB/s on these layers; the bulk store writes full lines and clips the
//     ragged last tile in hardware.
// Persistent CTAs (two per SM when the N tile is <= 128), warp roles: 0 = TMA producer, 1 = TMEM owner + MMA issuer,
// 2..5 = epilogue.
#include "umma.cuh"

namespace mtb {

using namespace um;

constexpr int PW_THREADS = 192;
constexpr int PW_RED_MAX16 = 4;  // fused reduction: N <= 64 = four 16-channel column groups
constexpr int PW_MAX_STAGES = 8;

struct PwParams {
  CUtensorMap a_map, w_map, o_map;
  const float* bias;
  double* stats;
  long long tiles_per_b;   // voxels per sample / 128 (statistics only)
  long long ntiles_rows;   // voxels in the whole tensor
  int ntiles;
  int KC, nkc, BN;         // channels per K chunk, chunks, N = padded Cout
  int OB, nob, out_mask;   // staging column block (channels), blocks, swizzle mask (7 / 3 / 1, 0 = dense)
  int widx;
  int stages, tmem_cols;
  int a_stage_bytes, w_chunk_bytes, out_buf_bytes;
  int obufs;               // staging buffers: 2, or 1 when two do not fit beside the weights
  int Cout_stride;         // stats row length (padded Cout)
  int is_f16;
  // fused InstanceNorm-backward reduction of the producing layer (see mtb200_conv_params::red): the statistics slots
  // then accumulate {sum dv, sum dv * xhat} instead of {sum x, sum x^2}
  const void* red_y;
  const float4* red_xform;
  const float2* red_meanrstd;
  int red_ldc, red_coff;
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

template <typename T>
__global__ void __launch_bounds__(PW_THREADS, 2) conv_pw_umma_kernel(const __grid_constant__ PwParams p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t full_bar[PW_MAX_STAGES], empty_bar[PW_MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2], w_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_sum[4][256], s_sq[4][256];  // statistics: one slot per epilogue warp, no float atomics
  __shared__ __align__(16) float s_bias[256];
  // fused reduction (N <= 64, no bias): {scale, shift, rstd, -mean * rstd} per channel ALIASES the bias table (two CTAs
  // per SM leave no room for another kilobyte of static shared memory), the LeakyReLU slopes get 256 bytes of their own
  float4* s_rc = reinterpret_cast<float4*>(s_bias);
  __shared__ float s_rslope[64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_base = dsmem;
  uint8_t* w_base = a_base + (size_t)p.stages * p.a_stage_bytes;
  uint8_t* o_base = w_base + (size_t)p.nkc * p.w_chunk_bytes;
  const uint32_t row_bytes = p.KC * 2;
  const int n0 = blockIdx.y * p.BN;  // N tile (Cout > 256 only)

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    mbar_init(&w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += PW_THREADS) {
#pragma unroll
    for (int w = 0; w < 4; ++w) { s_sum[w][i] = 0.f; s_sq[w][i] = 0.f; }
    s_bias[i] = (p.bias && i < p.BN) ? p.bias[n0 + i] : 0.f;
  }
  if (warp == 1) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: the weights once, then one box per (tile, chunk) =====
    if (elect_one()) {
      mbar_expect_tx(&w_full, (uint32_t)(p.nkc * p.BN * row_bytes));
      for (int c = 0; c < p.nkc; ++c)
        tma_load_3d(w_base + (size_t)c * p.w_chunk_bytes, &p.w_map, &w_full, c * p.KC, n0, p.widx);
    }
    __syncwarp();
    uint32_t gi = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      for (int c = 0; c < p.nkc; ++c, ++gi) {
        const uint32_t stage = gi % (uint32_t)p.stages;
        mbar_wait(&empty_bar[stage], ((gi / (uint32_t)p.stages) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], 128u * row_bytes);
          tma_load_2d(a_base + (size_t)stage * p.a_stage_bytes, &p.a_map, &full_bar[stage], c * p.KC, tile * 128);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
    const uint32_t idesc = idesc_f16(p.is_f16 != 0, (uint32_t)p.BN, false, false);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a16 = __shfl_sync(0xffffffffu, (smem_u32(a_base) & 0x3FFFFu) >> 4, 0);
    const uint32_t w16 = __shfl_sync(0xffffffffu, (smem_u32(w_base) & 0x3FFFFu) >> 4, 0);
    const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
    const uint32_t hi = (((8u * row_bytes) >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
    const uint32_t stage16 = (uint32_t)p.a_stage_bytes >> 4, wchunk16 = (uint32_t)p.w_chunk_bytes >> 4;
    const int ksteps = p.KC / 16;
    mbar_wait(&w_full, 0);
    tc_fence_after();
    uint32_t gi = 0, k = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++k) {
      const uint32_t buf = k & 1u;
      mbar_wait(&acc_empty[buf], ((k >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t dcol = tmem_u + buf * (uint32_t)p.BN;
      for (int c = 0; c < p.nkc; ++c, ++gi) {
        const uint32_t stage = gi % (uint32_t)p.stages;
        mbar_wait(&full_bar[stage], (gi / (uint32_t)p.stages) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = a16 + stage * stage16;
          const uint32_t sb = w16 + (uint32_t)c * wchunk16;
          for (int ks = 0; ks < ksteps; ++ks)
            umma_f16(dcol, ((uint64_t)hi << 32) | (uint64_t)(sa + 2u * ks), ((uint64_t)hi << 32) | (uint64_t)(sb + 2u * ks),
                     idesc, (c > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (c == p.nkc - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: warps 2..5; warp w owns TMEM lanes [32*(w%4), +32) = rows of the tile =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool want_stats = p.stats != nullptr;
    const bool red = want_stats && p.red_y != nullptr;
    const bool issuer = threadIdx.x == 64;
    const uint32_t ob_bytes = 128u * (uint32_t)p.OB * 2u;  // one staged column block
    uint32_t k = 0;
    int cur_b = -1;
    Raw8<T> yn[PW_RED_MAX16][2];
    if (red && (long long)blockIdx.x * 128 + row < (long long)p.ntiles_rows) {  // y of the first tile
      const T* yrow = reinterpret_cast<const T*>(p.red_y) + ((long long)blockIdx.x * 128 + row) * p.red_ldc + p.red_coff + n0;
#pragma unroll
      for (int i = 0; i < PW_RED_MAX16; ++i)
        if (i * 16 < p.BN) { yn[i][0].load(yrow + i * 16); yn[i][1].load(yrow + i * 16 + 8); }
    }
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++k) {
      const uint32_t buf = k & 1u;
      if (red) {  // constants of (sample, channel); a tile never straddles two samples (host check)
        const int b = (int)((long long)tile / p.tiles_per_b);
        if (b != cur_b) {
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int c = threadIdx.x - 64; c < p.BN; c += 128) {
            const long long i = (long long)b * p.Cout_stride + n0 + c;
            const float4 f = p.red_xform[i];
            const float2 mr = p.red_meanrstd[i];
            s_rc[c] = make_float4(f.x, f.y, mr.y, -mr.x * mr.y);
            s_rslope[c] = f.z;
          }
          cur_b = b;  // the bar.sync below orders these writes before their first use
        }
      }
      uint8_t* stage_out = o_base + (size_t)(p.obufs == 2 ? buf : 0u) * p.out_buf_bytes;
      // the bulk store that last read this staging buffer (tile k-2, or k-1) must have finished reading shared memory
      if (issuer) {
        if (p.obufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&acc_full[buf], (k >> 1) & 1u);
      tc_fence_after();
      const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)p.BN;
      // fused reduction: this tile's raw outputs y were requested while the PREVIOUS tile was processed (their global-load
      // latency would otherwise sit in the critical path of every tile); request the next tile's now
      const bool row_ok = (long long)tile * 128 + row < (long long)p.ntiles_rows;
      Raw8<T> yc[PW_RED_MAX16][2];
      if (red) {
#pragma unroll
        for (int i = 0; i < PW_RED_MAX16; ++i) { yc[i][0] = yn[i][0]; yc[i][1] = yn[i][1]; }
        const long long nrow = ((long long)tile + gridDim.x) * 128 + row;
        if (nrow < (long long)p.ntiles_rows) {
          const T* yrow = reinterpret_cast<const T*>(p.red_y) + nrow * p.red_ldc + p.red_coff + n0;
#pragma unroll
          for (int i = 0; i < PW_RED_MAX16; ++i)
            if (i * 16 < p.BN) { yn[i][0].load(yrow + i * 16); yn[i][1].load(yrow + i * 16 + 8); }
        }
      }
#pragma unroll 1
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t r[16];
        Raw8<T> y0, y1;
        if (red) {
#pragma unroll
          for (int i = 0; i < PW_RED_MAX16; ++i)
            if (i * 16 == c0) { y0 = yc[i][0]; y1 = yc[i][1]; }
        }
        tmem_ld16(tcol + (uint32_t)c0, r);
        float lo[8], hi8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          lo[j] = __uint_as_float(r[j]) + (red ? 0.f : s_bias[c0 + j]);
          hi8[j] = __uint_as_float(r[8 + j]) + (red ? 0.f : s_bias[c0 + 8 + j]);
        }
        const int blk = c0 / p.OB, cin = c0 - blk * p.OB;
        uint32_t off0 = (uint32_t)row * (uint32_t)p.OB * 2u + (uint32_t)cin * 2u;
        uint32_t off1 = off0 + 16u;
        off0 ^= ((off0 >> 7) & (uint32_t)p.out_mask) << 4;
        off1 ^= ((off1 >> 7) & (uint32_t)p.out_mask) << 4;
        uint8_t* ob = stage_out + (size_t)blk * ob_bytes;
        store8<T>(reinterpret_cast<T*>(ob + off0), lo);
        store8<T>(reinterpret_cast<T*>(ob + off1), hi8);
        if (want_stats) {
          float sv[16], ss[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x0 = Traits<T>::round(lo[j]), x1 = Traits<T>::round(hi8[j]);
            if (red) {  // {dv, dv * xhat} of the producing layer's InstanceNorm + LeakyReLU backward
              const float4 ca = s_rc[c0 + j], cb = s_rc[c0 + 8 + j];
              const float ya = row_ok ? y0.get(j) : 0.f, yb = row_ok ? y1.get(j) : 0.f;
              const float da = !row_ok ? 0.f : (fmaf(ya, ca.x, ca.y) > 0.f ? x0 : x0 * s_rslope[c0 + j]);
              const float db = !row_ok ? 0.f : (fmaf(yb, cb.x, cb.y) > 0.f ? x1 : x1 * s_rslope[c0 + 8 + j]);
              sv[j] = da; ss[j] = da * fmaf(ya, ca.z, ca.w);
              sv[8 + j] = db; ss[8 + j] = db * fmaf(yb, cb.z, cb.w);
            } else {
              sv[j] = x0; ss[j] = x0 * x0;
              sv[8 + j] = x1; ss[8 + j] = x1 * x1;
            }
          }
          warp_colsum16(sv, lane);
          warp_colsum16(ss, lane);
          if ((lane & 1) == 0) {
            const int col = colsum16_column(lane);
            s_sum[q][c0 + col] += sv[0];  // (warp, column) has exactly one owner lane
            s_sq[q][c0 + col] += ss[0];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      // staging tile complete: make the generic-proxy writes visible to the async proxy, then one thread stores
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (issuer) {
        for (int blk = 0; blk < p.nob; ++blk)
          tma_store_2d(&p.o_map, stage_out + (size_t)blk * ob_bytes, n0 + blk * p.OB, tile * 128);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (want_stats) {
        // flush the per-(b, channel) partials when this CTA moves on to another sample (host: nvox % 128 == 0 here)
        const long long next = (long long)tile + gridDim.x;
        if (next >= p.ntiles || next / p.tiles_per_b != (long long)tile / p.tiles_per_b) {
          const int b = (int)((long long)tile / p.tiles_per_b);
          for (int c = threadIdx.x - 64; c < p.BN; c += 128) {
            const float su = ((s_sum[0][c] + s_sum[1][c]) + s_sum[2][c]) + s_sum[3][c];
            const float sq = ((s_sq[0][c] + s_sq[1][c]) + s_sq[2][c]) + s_sq[3][c];
            if (su != 0.f || sq != 0.f) {
              double* st = p.stats + ((long long)b * p.Cout_stride + n0 + c) * 2;
              atomicAdd(st, (double)su);
              atomicAdd(st + 1, (double)sq);
#pragma unroll
              for (int w = 0; w < 4; ++w) { s_sum[w][c] = 0.f; s_sq[w][c] = 0.f; }
            }
          }
          // the next tile's first bar.sync orders these resets before any further accumulation
        }
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

static inline int pw_align1k(long long v) { return (int)(((v + 1023) / 1024) * 1024); }

// Returns MTB200_ERR_UNSUPPORTED when the problem is outside this kernel's envelope (caller falls back).
int conv_pw_umma(const mtb200_conv_params& p, cudaStream_t s) {
  if (p.ngroups != 1 || p.ntaps != 1 || p.accumulate || p.xform) return MTB200_ERR_UNSUPPORTED;
  for (int k = 0; k < 3; ++k)
    if (p.is[k] != 1 || p.os[k] != 1 || p.group_ooff[0][k] != 0 || p.tap_off[0][k] != 0) return MTB200_ERR_UNSUPPORTED;
  if (p.Do != p.Di || p.Ho != p.Hi || p.Wo != p.Wi || p.Dof != p.Do || p.Hof != p.Ho || p.Wof != p.Wo)
    return MTB200_ERR_UNSUPPORTED;
  if (p.Cout > 512 || p.Cout % 16) return MTB200_ERR_UNSUPPORTED;
  const long long nvox = (long long)p.Do * p.Ho * p.Wo;
  const long long M = nvox * p.B;
  if (M >= (1LL << 31) - 256) return MTB200_ERR_UNSUPPORTED;
  if (p.stats && nvox % 128) return MTB200_ERR_UNSUPPORTED;  // a tile must not straddle two samples

  static thread_local PwParams q;
  memset(&q, 0, sizeof(q));
  q.KC = p.Cin <= 16 ? 16 : (p.Cin <= 32 ? 32 : 64);
  q.nkc = (p.Cin + q.KC - 1) / q.KC;
  q.BN = p.Cout;
  if (q.BN > 256) {  // N tiles (grid.y); every tile re-streams the activations
    q.BN = 0;
    for (int c = 256; c >= 16; c -= 16)
      if (p.Cout % c == 0) { q.BN = c; break; }
  }
  const int rowb = q.KC * 2;
  const int out_rowb = q.BN * 2;
  if (out_rowb == 32 || out_rowb == 64 || out_rowb == 128) {
    q.OB = q.BN; q.out_mask = out_rowb == 128 ? 7 : (out_rowb == 64 ? 3 : 1);
  } else if (q.BN % 64 == 0) {
    q.OB = 64; q.out_mask = 7;
  } else {
    q.OB = q.BN; q.out_mask = 0;  // dense rows (e.g. 48 channels = 96 bytes), no swizzle
  }
  q.nob = q.BN / q.OB;
  q.a_stage_bytes = 128 * rowb;
  q.w_chunk_bytes = pw_align1k((long long)q.BN * rowb);
  q.out_buf_bytes = pw_align1k(128LL * out_rowb);
  q.tmem_cols = 32;
  while (q.tmem_cols < 2 * q.BN) q.tmem_cols *= 2;
  // two CTAs per SM when TMEM and shared memory allow it (small N, small weights), else one
  int per_sm = q.tmem_cols <= 256 ? 2 : 1;
  int fixed = 0;
  for (;; per_sm = 1) {
    const int budget = (per_sm == 2 ? 102 : 212) * 1024;  // + 10.5 KB static (per-warp statistics slots, bias) + 1 KB reserved per CTA
    q.obufs = 2;
    fixed = q.nkc * q.w_chunk_bytes + 2 * q.out_buf_bytes + 1024;
    if ((budget - fixed) / q.a_stage_bytes < 3) {
      q.obufs = 1;
      fixed -= q.out_buf_bytes;
    }
    q.stages = min(PW_MAX_STAGES, (budget - fixed) / q.a_stage_bytes);
    if (q.stages >= 2 || per_sm == 1) break;
  }
  if (q.stages < 2) return MTB200_ERR_UNSUPPORTED;
  q.ntiles = (int)((M + 127) / 128);
  q.tiles_per_b = max(1LL, nvox / 128);
  q.widx = p.tap_widx[0];
  q.bias = p.bias; q.stats = p.stats;
  q.ntiles_rows = M;
  // fused reduction of the producing layer's InstanceNorm backward (data-gradient launches; MTB200_FUSE_RED=0: off)
  // Off by default: the per-tile transposing reduction makes this HBM-streaming kernel latency-bound (head data gradient
  // at 192x160x128: 0.41 -> 1.45 ms, profiles/r2d_ncu_red.txt), more than the separate pass costs (0.38 ms).
  const char* fuse_env = getenv("MTB200_FUSE_RED_PW");  // read per call: tests toggle it
  const bool fuse_red = fuse_env && atoi(fuse_env) == 1;
  const bool red = p.red && fuse_red && !p.stats && !p.bias && p.red_y && p.red_xform && p.red_meanrstd && q.BN <= 64 &&
                   p.Cout == q.BN && nvox % 128 == 0 && p.red_ldc % 8 == 0 && p.red_coff % 8 == 0;
  if (red) {
    q.stats = p.red;  // the statistics slots and their flush carry {sum dv, sum dv * xhat}
    q.red_y = p.red_y;
    q.red_xform = reinterpret_cast<const float4*>(p.red_xform);
    q.red_meanrstd = reinterpret_cast<const float2*>(p.red_meanrstd);
    q.red_ldc = p.red_ldc; q.red_coff = p.red_coff;
  }
  q.Cout_stride = p.Cout;
  q.is_f16 = p.dtype == MTB200_F16;
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)p.in_ldc * 2};
    cuuint32_t box[2] = {(cuuint32_t)q.KC, 128};
    if (!umma_encode_map(&q.a_map, p.dtype, 2, (uint8_t*)p.in + (size_t)p.in_coff * 2, dims, strides, box, rowb))
      return MTB200_ERR_CUDA;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, (cuuint64_t)(q.widx + 1)};
    cuuint64_t strides[2] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)p.Cin * p.Cout * 2};
    cuuint32_t box[3] = {(cuuint32_t)q.KC, (cuuint32_t)q.BN, 1};
    if (!umma_encode_map(&q.w_map, p.dtype, 3, (void*)p.w, dims, strides, box, rowb)) return MTB200_ERR_CUDA;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.Cout, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)p.out_ldc * 2};
    cuuint32_t box[2] = {(cuuint32_t)q.OB, 128};
    if (!umma_encode_map(&q.o_map, p.dtype, 2, (uint8_t*)p.out + (size_t)p.out_coff * 2, dims, strides, box,
                         q.out_mask ? q.OB * 2 : 0))
      return MTB200_ERR_CUDA;
  }
  const int smem = q.stages * q.a_stage_bytes + fixed;
  const int ny = p.Cout / q.BN;
  const int gx = (int)min((long long)q.ntiles, max(1LL, (long long)per_sm * num_sms() / ny));
  dim3 grid((unsigned)gx, (unsigned)ny, 1);
  cudaError_t e;
  if (p.dtype == MTB200_BF16) {
    e = cudaFuncSetAttribute(conv_pw_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) launch_pdl(conv_pw_umma_kernel<__nv_bfloat16>, dim3(grid), dim3(PW_THREADS), (size_t)(smem), s, q);
  } else {
    e = cudaFuncSetAttribute(conv_pw_umma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) launch_pdl(conv_pw_umma_kernel<__half>, dim3(grid), dim3(PW_THREADS), (size_t)(smem), s, q);
  }
  if (e != cudaSuccess) { set_error("conv_pw: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  return check_launch(q.red_y ? "conv_pw_umma+red" : "conv_pw_umma");
}

}  // namespace mtb
