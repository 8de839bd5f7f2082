// First layer of the U-Net: Conv3d(1 -> 30, 3x3x3, stride 1, padding 1) on the single-channel CT patch
// (generic_UNet.py:46 with input_channels = 1; conv_blocks_context.0.blocks.0) and its weight gradient.
//
// With one input channel the generic kernels pad Cin to 16 and spend 9 (forward) / 24 (wgrad) N = 96 MMAs per 128 voxels
// on a layer that does 54 FLOP per output byte -- it is an HBM stream (1 GB of bf16 output per bs4 step).  Here the GEMM
// K dimension is the TAP index instead: four builder warps assemble the im2col tile A[128 voxels][32 = 27 taps + 5 zeros]
// (64-byte rows, SWIZZLE_64B image) straight from the 1-channel volume (27 two-byte loads per voxel, L1-resident), and
//   forward : D[128 vox][Cout] = A . B^T,  B[co][tap] resident (K-major), 2 MMAs per tile, pointwise-style epilogue
//             (+bias, InstanceNorm sum / sum-of-squares, staged TMA store);
//   wgrad   : D[tap][co] += A^T . dY -- the SAME shared-memory image read as an MN-major operand (rows = K = voxels),
//             dY tiles by TMA, one TMEM accumulator per CTA for the whole kernel, 27 x Cout fp32 atomics at the end.
// Persistent CTAs (several per SM), work unit = (b, d, h, 128-wide w tile).
#include "umma.cuh"

namespace mtb {

using namespace um;

constexpr int C1_STAGES = 4;
constexpr int C1_FWD_STAGES = 6;
constexpr int C1_FWD_BUILDERS = 8;     // two groups of four builder warps take tiles round-robin
constexpr int C1_FWD_EPI = 16;         // two groups of eight epilogue warps take alternate tiles (= the two TMEM buffers)
constexpr int C1_FWD_THREADS = 32 * (C1_FWD_BUILDERS + 1 + C1_FWD_EPI);
constexpr int C1_WG_THREADS = 448;     // 8 builder warps, MMA warp, dY producer warp, 4 epilogue warps
constexpr int C1_ROWB = 64;            // bytes per im2col row (32 taps x 2 B)

struct C1Params {
  CUtensorMap o_map;       // forward: output / wgrad: dY, both {C, W, B*D*H} with box {C, 128, 1}
  const void* x;           // single-channel volume, element stride xs between consecutive voxels
  long long xs;
  const void* w;           // packed [27][Cout][Cin_p] (only ci = 0 is read)
  int Cin_p;
  const float* bias;
  double* stats;
  float* dw;               // wgrad: [27][Cout][Cin_p] fp32
  int B, D, H, W, Cout;
  int ntw;
  uint32_t units;
  int out_mask;            // staging swizzle mask of the Cout * 2-byte rows
  int is_f16;
};

// one im2col row per thread: 27 taps of voxel (b, d, h, w) -> four swizzled 16-byte chunks of stage row r.
// COMPACT (voxel stride 1): the three dx taps of a line are immediate offsets of one pointer.
template <bool COMPACT>
__device__ __forceinline__ void c1_build_row(const C1Params& p, uint8_t* stage, int r, int b, int d, int h, int w) {
  const uint16_t* x = reinterpret_cast<const uint16_t*>(p.x);
  const long long xs = COMPACT ? 1 : p.xs;
  const bool okl = w >= 1 && w <= p.W, okc = w < p.W, okr = w + 1 < p.W;
  const uint16_t* center = x + ((((long long)b * p.D + d) * p.H + h) * p.W + w) * xs;
  const long long sh = (long long)p.W * xs, sd = (long long)p.H * sh;
  uint32_t v[27];
#pragma unroll
  for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const bool line_ok = (unsigned)(d + dz) < (unsigned)p.D && (unsigned)(h + dy) < (unsigned)p.H;
      const uint16_t* line = center + dz * sd + dy * sh;
      const int t = ((dz + 1) * 3 + (dy + 1)) * 3;
      v[t] = 0u; v[t + 1] = 0u; v[t + 2] = 0u;
      if (line_ok && okl) v[t] = (uint32_t)__ldg(line - xs);
      if (line_ok && okc) v[t + 1] = (uint32_t)__ldg(line);
      if (line_ok && okr) v[t + 2] = (uint32_t)__ldg(line + xs);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t wd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t0 = c * 8 + 2 * j, t1 = t0 + 1;
      const uint32_t lo = t0 < 27 ? v[t0 < 27 ? t0 : 0] : 0u, hi = t1 < 27 ? v[t1 < 27 ? t1 : 0] : 0u;
      wd[j] = lo | (hi << 16);
    }
    uint32_t off = (uint32_t)r * C1_ROWB + (uint32_t)c * 16u;
    off ^= ((off >> 7) & 3u) << 4;
    *reinterpret_cast<uint4*>(stage + off) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
  }
}

// work unit -> (b, d, h, w tile); 32-bit arithmetic (64-bit divisions here cost more than the tile's real work)
__device__ __forceinline__ void c1_unit(const C1Params& p, uint32_t u, int& b, int& d, int& h, int& w0) {
  const uint32_t ntw = (uint32_t)p.ntw, H = (uint32_t)p.H, D = (uint32_t)p.D;
  const uint32_t tw = ntw == 1 ? 0u : u % ntw;
  uint32_t row = ntw == 1 ? u : u / ntw;
  const uint32_t rh = row / H;
  h = (int)(row - rh * H);
  const uint32_t rd = rh / D;
  d = (int)(rh - rd * D);
  b = (int)rd;
  w0 = (int)tw * 128;
}

// (b, d, h, w tile) of a work unit, advanced by a FIXED unit stride with carries instead of divisions: the three 32-bit
// divisions of c1_unit per tile in every builder and epilogue thread were a third of the forward kernel's instructions
struct C1Walk {
  int b, d, h, tw;
  int sb, sd, sh, stw;
  __device__ __forceinline__ void init(const C1Params& p, uint32_t u, uint32_t stride) {
    int w0;
    c1_unit(p, u, b, d, h, w0);
    tw = w0 >> 7;
    const uint32_t ntw = (uint32_t)p.ntw, H = (uint32_t)p.H, D = (uint32_t)p.D;
    stw = (int)(stride % ntw);
    uint32_t r = stride / ntw;
    sh = (int)(r % H); r /= H;
    sd = (int)(r % D);
    sb = (int)(r / D);
  }
  __device__ __forceinline__ void next(const C1Params& p) {
    tw += stw; h += sh; d += sd; b += sb;
    if (tw >= p.ntw) { tw -= p.ntw; ++h; }
    if (h >= p.H) { h -= p.H; ++d; }
    if (d >= p.D) { d -= p.D; ++b; }
  }
};

__device__ __forceinline__ void tma_load_3d_tile(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  tma_load_3d(dst, map, bar, c0, c1, c2);
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// =====================================================================================================================
// forward
// =====================================================================================================================
// One CTA per SM.  The epilogue was the critical path of the first version (4 warps, a 32-lane transposing butterfly and
// shared-memory atomics per tile for the InstanceNorm statistics: the builders waited 40 % of the time for free stages,
// profiles/r1i_ncu_full_taps_c1.txt): now 8 epilogue warps (two per TMEM lane quarter, each half of the channels) keep
// the per-(b, channel) sums in REGISTERS across tiles and reduce them once per sample.  NCB = 16-channel blocks per
// epilogue warp (Cout = 32 * NCB).
//
// LINES (compact volumes): K = the nine (dz, dy) input LINES instead of the 27 taps.  A builder thread owns one voxel column
// w of the tile's 130 (halo included) and writes ONE 32-byte row A[w][k] = x[d+dz][h+dy][w] (9 two-byte loads, zeros for
// k = 9..15); the three dx taps are ROW-SHIFTED views of that tile (start row dx + 1), i.e. 3 MMAs of K = 16 against
// B_dx[co][9 lines].  A third of the loads and of the packing work of the 27-tap im2col (which limited the first version to
// 0.74 ms for a 0.17 ms HBM stream).  (Fetching the lines by TMA and shifting by coordinates is not possible: a one-voxel
// shift of a single-channel line is a 2-byte offset, and bulk tensor copies start on 16-byte boundaries.)
template <typename T, int NCB, bool LINES>
__global__ void __launch_bounds__(C1_FWD_THREADS, 1) conv_c1_fwd_kernel(const __grid_constant__ C1Params p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C1_FWD_STAGES], empty_bar[C1_FWD_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[64];

  constexpr int MMA_WARP = C1_FWD_BUILDERS;
  constexpr int EPI_WARP0 = C1_FWD_BUILDERS + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  const int a_stage_bytes = LINES ? 5120 : 128 * C1_ROWB;            // 8 KB im2col tile / 130 (+ slack) rows of 32 bytes
  uint8_t* a_base = dsmem;
  uint8_t* b_tile = a_base + C1_FWD_STAGES * a_stage_bytes;          // [Cout][64 B], K-major, SWIZZLE_64B image
  const int out_buf_bytes = ((128 * p.Cout * 2 + 1023) / 1024) * 1024;
  uint8_t* o_base = b_tile + (LINES ? 8192 : 4096);                  // LINES: three [Cout][32 B] tiles (one per dx)
  const uint32_t tmem_cols = (uint32_t)(2 * p.Cout);

  if (threadIdx.x == 0) {
    for (int i = 0; i < C1_FWD_STAGES; ++i) { mbar_init(&full_bar[i], 4); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 64) s_bias[threadIdx.x] = (p.bias && (int)threadIdx.x < p.Cout) ? p.bias[threadIdx.x] : 0.f;
  if (LINES) {
    // B_dx[co][k = (dz, dy)] (32-byte rows, SWIZZLE_32B image), zero for k = 9..15; and the rows 9..15 of every line box
    for (int i = threadIdx.x; i < 3 * p.Cout * 2; i += blockDim.x) {
      const int dx = i / (p.Cout * 2), co = (i >> 1) % p.Cout, c = i & 1;
      const uint16_t* wp = reinterpret_cast<const uint16_t*>(p.w);
      uint32_t wd[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k0 = c * 8 + 2 * j, k1 = k0 + 1;
        const uint32_t lo = k0 < 9 ? (uint32_t)wp[((long long)(k0 * 3 + dx) * p.Cout + co) * p.Cin_p] : 0u;
        const uint32_t hi = k1 < 9 ? (uint32_t)wp[((long long)(k1 * 3 + dx) * p.Cout + co) * p.Cin_p] : 0u;
        wd[j] = lo | (hi << 16);
      }
      *reinterpret_cast<uint4*>(b_tile + dx * 2048 + co * 32 + ((c ^ ((co >> 2) & 1)) * 16)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
  }
  // weight tile B[co][tap] from the packed weights (ci = 0), zero for taps 27..31
  for (int i = threadIdx.x; !LINES && i < p.Cout * 4; i += blockDim.x) {
    const int co = i >> 2, c = i & 3;
    const uint16_t* wp = reinterpret_cast<const uint16_t*>(p.w);
    uint32_t wd[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t0 = c * 8 + 2 * j, t1 = t0 + 1;
      const uint32_t lo = t0 < 27 ? (uint32_t)wp[((long long)t0 * p.Cout + co) * p.Cin_p] : 0u;
      const uint32_t hi = t1 < 27 ? (uint32_t)wp[((long long)t1 * p.Cout + co) * p.Cin_p] : 0u;
      wd[j] = lo | (hi << 16);
    }
    uint32_t off = (uint32_t)co * C1_ROWB + (uint32_t)c * 16u;
    off ^= ((off >> 7) & 3u) << 4;
    *reinterpret_cast<uint4*>(b_tile + off) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == MMA_WARP) tmem_alloc(&tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (LINES && warp < C1_FWD_BUILDERS) {
    // ===== line-tile builders: thread = one voxel column of the tile (130 with the halo: threads 0, 1 take a second row) =====
    constexpr uint32_t NG = C1_FWD_BUILDERS / 4;
    const uint32_t grp = (uint32_t)warp >> 2;
    const int r0 = (warp & 3) * 32 + lane;
    const uint16_t* xv = reinterpret_cast<const uint16_t*>(p.x);
    const long long sh = p.W, sd = (long long)p.H * p.W;
    // the nine line values of tile row r (voxel column w0 - 1 + r) of work unit u
    auto fetch = [&](const C1Walk& k, int r, uint32_t (&v)[9]) {
      const int b = k.b, d = k.d, h = k.h, w0 = k.tw * 128;
      const int w = w0 - 1 + r;
      const bool okw = (unsigned)w < (unsigned)p.W;
      const uint16_t* plane = xv + (((long long)b * p.D + d) * p.H + h) * p.W;
#pragma unroll
      for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const bool ok = okw && (unsigned)(d + dz) < (unsigned)p.D && (unsigned)(h + dy) < (unsigned)p.H;
          v[(dz + 1) * 3 + (dy + 1)] = ok ? (uint32_t)__ldg(plane + dz * sd + dy * sh + w) : 0u;
        }
    };
    auto put = [&](uint8_t* st, int r, const uint32_t (&v)[9]) {
      const uint4 c0 = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
      const uint4 c1 = make_uint4(v[8], 0u, 0u, 0u);
      const int sw = (r >> 2) & 1;  // SWIZZLE_32B image: 16-byte pieces of rows 4..7 of every 8 swapped
      *reinterpret_cast<uint4*>(st + r * 32 + (sw ? 16 : 0)) = c0;
      *reinterpret_cast<uint4*>(st + r * 32 + (sw ? 0 : 16)) = c1;
    };
    // software pipeline: the loads of this group's NEXT tile are in flight while the current one is packed and handed over
    // (a tile's lines come from L2 / HBM: their latency was the critical path of the whole kernel)
    uint32_t va[9], vb[9], na[9], nb[9];
    const uint32_t ustep = NG * gridDim.x;
    uint32_t u = blockIdx.x + grp * gridDim.x;
    C1Walk wk;
    wk.init(p, u < p.units ? u : 0u, ustep);
    if (u < p.units) {
      fetch(wk, r0, na);
      if (r0 < 2) fetch(wk, r0 + 128, nb);
    }
    for (uint32_t gi = grp; u < p.units; gi += NG, u += ustep) {
#pragma unroll
      for (int k = 0; k < 9; ++k) { va[k] = na[k]; vb[k] = nb[k]; }
      if (u + ustep < p.units) {
        wk.next(p);
        fetch(wk, r0, na);
        if (r0 < 2) fetch(wk, r0 + 128, nb);
      }
      const uint32_t stage = gi % C1_FWD_STAGES;
      mbar_wait(&empty_bar[stage], ((gi / C1_FWD_STAGES) & 1u) ^ 1u);
      uint8_t* st = a_base + (size_t)stage * a_stage_bytes;
      put(st, r0, va);
      if (r0 < 2) put(st, r0 + 128, vb);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    }
  } else if (warp < C1_FWD_BUILDERS) {
    // ===== im2col builders: thread = one voxel of the tile =====
    // three groups of four builder warps take tiles round-robin: one group's load latency (a new input line from L2 /
    // HBM per tile, exposed by the proxy fence before the arrive) overlaps the other groups' packing
    constexpr uint32_t NG = C1_FWD_BUILDERS / 4;
    const uint32_t grp = (uint32_t)warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    for (uint32_t gi = grp, u = blockIdx.x + grp * gridDim.x; u < p.units; gi += NG, u += NG * gridDim.x) {
      int b, d, h, w0;
      c1_unit(p, u, b, d, h, w0);
      const uint32_t stage = gi % C1_FWD_STAGES;
      mbar_wait(&empty_bar[stage], ((gi / C1_FWD_STAGES) & 1u) ^ 1u);
      uint8_t* st = a_base + (size_t)stage * a_stage_bytes;
      if (p.xs == 1) c1_build_row<true>(p, st, r, b, d, h, w0 + r);
      else c1_build_row<false>(p, st, r, b, d, h, w0 + r);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    }
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer =====
    const uint32_t idesc = idesc_f16(p.is_f16 != 0, (uint32_t)p.Cout, false, false);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a16 = __shfl_sync(0xffffffffu, (smem_u32(a_base) & 0x3FFFFu) >> 4, 0);
    const uint32_t b16 = __shfl_sync(0xffffffffu, (smem_u32(b_tile) & 0x3FFFFu) >> 4, 0);
    const uint32_t hi = (((8u * C1_ROWB) >> 4) & 0x3FFFu) | (1u << 14) | (4u << 29);  // SWIZZLE_64B
    // LINES: A and B K-major with 32-byte rows (SWIZZLE_32B, SBO = 8 rows); a dx tap = A from row dx + 1 on
    const uint32_t hi_b = ((8u * 32u) >> 4) | (1u << 14) | (6u << 29);
    uint32_t gi = 0;
    for (uint32_t u = blockIdx.x; u < p.units; u += gridDim.x, ++gi) {
      const uint32_t stage = gi % C1_FWD_STAGES, buf = gi & 1u;
      mbar_wait(&acc_empty[buf], ((gi >> 1) & 1u) ^ 1u);
      mbar_wait(&full_bar[stage], (gi / C1_FWD_STAGES) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = a16 + stage * (uint32_t)(a_stage_bytes >> 4);
        const uint32_t dcol = tmem_u + buf * (uint32_t)p.Cout;
        if (LINES) {
#pragma unroll
          for (int j = 0; j < 3; ++j)  // dx = j - 1: rows j .. j + 127 of the line tile against B_dx
            umma_f16(dcol, ((uint64_t)hi_b << 32) | (uint64_t)(sa + (uint32_t)(j * 2)),
                     ((uint64_t)hi_b << 32) | (uint64_t)(b16 + (uint32_t)(j * 128)), idesc, j ? 1u : 0u);
        } else {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            umma_f16(dcol, ((uint64_t)hi << 32) | (uint64_t)(sa + 2u * ks), ((uint64_t)hi << 32) | (uint64_t)(b16 + 2u * ks),
                     idesc, ks ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: two groups of 8 warps (group = tile parity = TMEM buffer = staging buffer: the per-tile chain of
    // TMEM load -> arithmetic -> staging -> barrier -> bulk store is latency, so two tiles are drained at the same time);
    // inside a group: TMEM lane quarter = warp % 4, channel half = (warp in group) / 4 =====
    const int eg = (warp - EPI_WARP0) >> 3;
    const int q = warp & 3;
    const int half = ((warp - EPI_WARP0) & 7) >> 2;
    const int row = q * 32 + lane;
    const int cbase = half * (16 * NCB);
    const bool want_stats = p.stats != nullptr;
    const bool issuer = ((warp - EPI_WARP0) & 7) == 0 && lane == 0;
    const int bar_id = 1 + eg;
    float csum[NCB][16], csq[NCB][16];
#pragma unroll
    for (int i = 0; i < NCB; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) { csum[i][j] = 0.f; csq[i][j] = 0.f; }
    auto flush = [&](int b) {  // per-(b, channel) sums of this warp's rows -> fp64 atomics
#pragma unroll
      for (int i = 0; i < NCB; ++i) {
        warp_colsum16(csum[i], lane);
        warp_colsum16(csq[i], lane);
        if ((lane & 1) == 0) {
          double* st = p.stats + ((long long)b * p.Cout + cbase + i * 16 + colsum16_column(lane)) * 2;
          atomicAdd(st, (double)csum[i][0]);
          atomicAdd(st + 1, (double)csq[i][0]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) { csum[i][j] = 0.f; csq[i][j] = 0.f; }
      }
    };
    int cur_b = -1;
    C1Walk wk;
    {
      const uint32_t u0 = blockIdx.x + (uint32_t)eg * gridDim.x;
      wk.init(p, u0 < p.units ? u0 : 0u, 2 * gridDim.x);
    }
    bool first_tile = true;
    for (uint32_t gi = (uint32_t)eg, u = blockIdx.x + (uint32_t)eg * gridDim.x; u < p.units; gi += 2, u += 2 * gridDim.x) {
      if (!first_tile) wk.next(p);
      first_tile = false;
      const int b = wk.b, d = wk.d, h = wk.h, w0 = wk.tw * 128;
      if (want_stats && b != cur_b) {
        if (cur_b >= 0) flush(cur_b);
        cur_b = b;
      }
      const uint32_t buf = gi & 1u;
      uint8_t* stage_out = o_base + (size_t)(eg * 2 + ((gi >> 1) & 1u)) * out_buf_bytes;  // two staging buffers per group
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // this group's store before the last
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      mbar_wait(&acc_full[buf], (gi >> 1) & 1u);
      tc_fence_after();
      const bool valid = w0 + row < p.W;
      const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)p.Cout;
#pragma unroll
      for (int i = 0; i < NCB; ++i) {
        const int c0 = cbase + i * 16;
        uint32_t r[16];
        tmem_ld16(tcol + (uint32_t)c0, r);
        float lo[8], hi8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          lo[j] = __uint_as_float(r[j]) + s_bias[c0 + j];
          hi8[j] = __uint_as_float(r[8 + j]) + s_bias[c0 + 8 + j];
        }
        uint32_t off0 = (uint32_t)row * (uint32_t)p.Cout * 2u + (uint32_t)c0 * 2u;
        uint32_t off1 = off0 + 16u;
        off0 ^= ((off0 >> 7) & (uint32_t)p.out_mask) << 4;
        off1 ^= ((off1 >> 7) & (uint32_t)p.out_mask) << 4;
        store8<T>(reinterpret_cast<T*>(stage_out + off0), lo);
        store8<T>(reinterpret_cast<T*>(stage_out + off1), hi8);
        if (want_stats && valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x0 = Traits<T>::round(lo[j]), x1 = Traits<T>::round(hi8[j]);
            csum[i][j] += x0; csq[i][j] = fmaf(x0, x0, csq[i][j]);
            csum[i][8 + j] += x1; csq[i][8 + j] = fmaf(x1, x1, csq[i][8 + j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      if (issuer) {
        tma_store_3d(&p.o_map, stage_out, 0, w0, (b * p.D + d) * p.H + h);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (want_stats && cur_b >= 0) flush(cur_b);
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// =====================================================================================================================
// weight gradient
// =====================================================================================================================
__global__ void __launch_bounds__(C1_WG_THREADS, 2) conv_c1_wgrad_kernel(const __grid_constant__ C1Params p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t a_full[C1_STAGES], y_full[C1_STAGES], empty_bar[C1_STAGES];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  const int a_stage_bytes = 128 * C1_ROWB;               // 8 KB: [128 voxels][32 taps]
  const int y_stage_bytes = 128 * p.Cout * 2;            // [128 voxels][Cout]
  uint8_t* a_base = dsmem;
  uint8_t* y_base = a_base + C1_STAGES * a_stage_bytes + 1024;  // slack: the shifted (unused) M blocks read 3 rows past a stage
  const bool have_work = blockIdx.x < p.units;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C1_STAGES; ++i) { mbar_init(&a_full[i], 4); mbar_init(&y_full[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(&tmem_slot, 64u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp < 8) {
    // two groups of four builder warps take alternate tiles: one group's load latency (a new input line from L2 / HBM
    // per tile, exposed by the proxy fence before the arrive) overlaps the other group's packing
    const int grp = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    for (uint32_t gi = (uint32_t)grp, u = blockIdx.x + (uint32_t)grp * gridDim.x; u < p.units; gi += 2, u += 2 * gridDim.x) {
      int b, d, h, w0;
      c1_unit(p, u, b, d, h, w0);
      const uint32_t stage = gi % C1_STAGES;
      mbar_wait(&empty_bar[stage], ((gi / C1_STAGES) & 1u) ^ 1u);
      uint8_t* st = a_base + (size_t)stage * a_stage_bytes;
      if (p.xs == 1) c1_build_row<true>(p, st, r, b, d, h, w0 + r);
      else c1_build_row<false>(p, st, r, b, d, h, w0 + r);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[stage]);
    }
  } else if (warp == 9) {
    // ===== dY producer (rows past the end of a line are zero-filled by TMA) =====
    uint32_t gi = 0;
    for (uint32_t u = blockIdx.x; u < p.units; u += gridDim.x, ++gi) {
      const uint32_t stage = gi % C1_STAGES;
      mbar_wait(&empty_bar[stage], ((gi / C1_STAGES) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&y_full[stage], (uint32_t)y_stage_bytes);
        tma_load_3d_tile(y_base + (size_t)stage * y_stage_bytes, &p.o_map, &y_full[stage], 0,
                         (int)(p.ntw == 1 ? 0u : u % (uint32_t)p.ntw) * 128, (int)(p.ntw == 1 ? u : u / (uint32_t)p.ntw));
      }
      __syncwarp();
    }
  } else if (warp == 8) {
    // ===== MMA issuer: D[128 = (shift block, tap)][Cout] += A^T . dY, both operands MN-major =====
    const uint32_t idesc = idesc_f16(p.is_f16 != 0, (uint32_t)p.Cout, true, true);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a16 = __shfl_sync(0xffffffffu, (smem_u32(a_base) & 0x3FFFFu) >> 4, 0);
    const uint32_t y16 = __shfl_sync(0xffffffffu, (smem_u32(y_base) & 0x3FFFFu) >> 4, 0);
    const uint32_t yrow = (uint32_t)p.Cout * 2u;
    const uint32_t hi_a = ((8u * C1_ROWB) >> 4) | (1u << 14) | (4u << 29);                  // SBO = 8 rows, SWIZZLE_64B
    const uint32_t lbo_a = ((uint32_t)C1_ROWB >> 4) << 16;                                    // M blocks 1..3: shifted views (unused)
    const uint32_t lay_y = yrow == 128 ? 2u : (yrow == 64 ? 4u : 6u);
    const uint32_t hi_y = ((8u * yrow) >> 4) | (1u << 14) | (lay_y << 29);
    uint32_t gi = 0;
    for (uint32_t u = blockIdx.x; u < p.units; u += gridDim.x, ++gi) {
      const uint32_t stage = gi % C1_STAGES;
      mbar_wait(&a_full[stage], (gi / C1_STAGES) & 1u);
      mbar_wait(&y_full[stage], (gi / C1_STAGES) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = (a16 + stage * (uint32_t)(a_stage_bytes >> 4)) | lbo_a;
        const uint32_t sy = y16 + stage * (uint32_t)(y_stage_bytes >> 4);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16(tmem_u, ((uint64_t)hi_a << 32) | (uint64_t)(sa + (uint32_t)(kk * C1_ROWB)),
                   ((uint64_t)hi_y << 32) | (uint64_t)(sy + (uint32_t)kk * yrow), idesc, (gi > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
    }
    if (have_work) {
      if (elect_one()) umma_commit(&acc_full);
      __syncwarp();
    }
  } else if (have_work) {
    // ===== epilogue (once): rows 0..26 of the accumulator = taps =====
    const int q = warp & 3;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    if (q == 0) {
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + (uint32_t)c0, r);
        if (lane < 27) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float v = __uint_as_float(r[j]);
            if (v != 0.f) atomicAdd(p.dw + ((long long)lane * p.Cout + c0 + j) * p.Cin_p, v);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64u);
  }
}

// =====================================================================================================================
// host
// =====================================================================================================================
static int c1_common(C1Params& q, const void* x, long long xs, int dtype, int B, int D, int H, int W, int Cout_p,
                     void* mat, int ldc, int coff, bool swizzled_box) {
  if (dtype != MTB200_BF16 && dtype != MTB200_F16) { set_error("conv_c1: 16-bit activations only"); return MTB200_ERR_UNSUPPORTED; }
  if (!umma_encode_fn()) { set_error("conv_c1: no tensor-map encoder (not an sm_100 driver)"); return MTB200_ERR_UNSUPPORTED; }
  if (Cout_p != 16 && Cout_p != 32 && Cout_p != 64) { set_error("conv_c1: Cout_p must be 16, 32 or 64"); return MTB200_ERR_UNSUPPORTED; }
  if (ldc % 8 || coff % 8) { set_error("conv_c1: channel stride / offset must be multiples of 8"); return MTB200_ERR_UNSUPPORTED; }
  memset(&q, 0, sizeof(q));
  q.x = x; q.xs = xs;
  q.B = B; q.D = D; q.H = H; q.W = W; q.Cout = Cout_p;
  q.ntw = (W + 127) / 128;
  if ((long long)B * D * H * q.ntw >= (1LL << 31) - 65536) { set_error("conv_c1: too many lines"); return MTB200_ERR_UNSUPPORTED; }
  q.units = (uint32_t)((long long)B * D * H * q.ntw);
  const int rowb = Cout_p * 2;
  q.out_mask = rowb == 128 ? 7 : (rowb == 64 ? 3 : 1);
  q.is_f16 = dtype == MTB200_F16;
  cuuint64_t dims[3] = {(cuuint64_t)Cout_p, (cuuint64_t)W, (cuuint64_t)B * D * H};
  cuuint64_t strides[2] = {(cuuint64_t)ldc * 2, (cuuint64_t)W * ldc * 2};
  cuuint32_t box[3] = {(cuuint32_t)Cout_p, 128, 1};
  (void)swizzled_box;
  if (!umma_encode_map(&q.o_map, dtype, 3, (uint8_t*)mat + (size_t)coff * 2, dims, strides, box, rowb)) return MTB200_ERR_CUDA;
  return MTB200_OK;
}

template <typename T, int NCB>
static cudaError_t launch_c1_fwd(const C1Params& q, int gx, int smem, bool tma_in, cudaStream_t s) {
  cudaError_t e;
  if (tma_in) {
    e = cudaFuncSetAttribute(conv_c1_fwd_kernel<T, NCB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) launch_pdl(conv_c1_fwd_kernel<T, NCB, true>, dim3(gx), dim3(C1_FWD_THREADS), (size_t)(smem), s, q);
  } else {
    e = cudaFuncSetAttribute(conv_c1_fwd_kernel<T, NCB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) launch_pdl(conv_c1_fwd_kernel<T, NCB, false>, dim3(gx), dim3(C1_FWD_THREADS), (size_t)(smem), s, q);
  }
  return e;
}

int conv_c1_fwd(const void* x, long long xs, const void* w, int Cin_p, const float* bias, void* out, int out_ldc,
                int out_coff, int Cout_p, double* stats, int dtype, int B, int D, int H, int W, cudaStream_t s) {
  static C1Params q;
  if (int r = c1_common(q, x, xs, dtype, B, D, H, W, Cout_p, out, out_ldc, out_coff, true)) return r;
  if (q.units == 0) return MTB200_OK;
  q.w = w; q.Cin_p = Cin_p; q.bias = bias; q.stats = stats;
  if (Cout_p != 32 && Cout_p != 64) { set_error("conv_c1_fwd: Cout_p must be 32 or 64"); return MTB200_ERR_UNSUPPORTED; }
  const int out_buf = ((128 * Cout_p * 2 + 1023) / 1024) * 1024;
  // K = lines variant: compact volume (MTB200_C1_LINES=0: the 27-tap im2col kernel)
  static const int lines_on = [] { const char* e = getenv("MTB200_C1_LINES"); return e ? atoi(e) : 1; }();
  const bool tma_in = lines_on && xs == 1;
  const int smem = C1_FWD_STAGES * (tma_in ? 5120 : 128 * C1_ROWB) + (tma_in ? 8192 : 4096) + 4 * out_buf + 1024;
  const int gx = (int)min((long long)q.units, (long long)num_sms());  // one persistent CTA per SM
  cudaError_t e;
  if (dtype == MTB200_BF16) {
    e = Cout_p == 32 ? launch_c1_fwd<__nv_bfloat16, 1>(q, gx, smem, tma_in, s) : launch_c1_fwd<__nv_bfloat16, 2>(q, gx, smem, tma_in, s);
  } else {
    e = Cout_p == 32 ? launch_c1_fwd<__half, 1>(q, gx, smem, tma_in, s) : launch_c1_fwd<__half, 2>(q, gx, smem, tma_in, s);
  }
  if (e != cudaSuccess) { set_error("conv_c1_fwd: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  return check_launch("conv_c1_fwd");
}

int conv_c1_wgrad(const void* x, long long xs, const void* dy, int dy_ldc, int dy_coff, int Cout_p, float* dw, int Cin_p,
                  int dtype, int B, int D, int H, int W, cudaStream_t s) {
  static C1Params q;
  if (int r = c1_common(q, x, xs, dtype, B, D, H, W, Cout_p, const_cast<void*>(dy), dy_ldc, dy_coff, true)) return r;
  if (q.units == 0) return MTB200_OK;
  q.dw = dw; q.Cin_p = Cin_p;
  const int smem = C1_STAGES * (128 * C1_ROWB + 128 * Cout_p * 2) + 1024 + 1024;
  const int gx = (int)min((long long)q.units, (long long)2 * num_sms());
  cudaError_t e = cudaFuncSetAttribute(conv_c1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("conv_c1_wgrad: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  launch_pdl(conv_c1_wgrad_kernel, dim3(gx), dim3(C1_WG_THREADS), (size_t)(smem), s, q);
  return check_launch("conv_c1_wgrad");
}

}  // namespace mtb
