// Line-streaming tcgen05 convolution with the dy taps merged into the MMA's N dimension -- for the NARROW full-resolution
// layers of the U-Net (Cout_p tile = 32, Cin_p <= 64; generic_UNet.py:46 at 192x160x128), stride 1, taps in [-1,1]^3.
//
// Why: measured on the B200 (profiles/r1d_conv_bench.txt) one tcgen05.mma M=128 costs ~50 ns whatever N <= 128 is (the
// 128 x 16 A tile has to be fetched from shared memory), so an implicit GEMM with N = Cout = 32 cannot exceed ~17 % of
// the tensor peak.  Here the three dy taps of every (dz, dx) share ONE MMA with N = 3 x 32 = 96:
//     Q[h'][w][(dy, co)] = sum_{dz, dx, ci} X[d+dz][h'][w+dx][ci] * W[dz, dy, dx][ci][co]        (h' = an INPUT line)
//     out[d][h][w][co]   = Q[h-1][..][(dy=-1, co)] + Q[h][..][(dy=0, co)] + Q[h+1][..][(dy=+1, co)]
// i.e. 3x fewer MMAs; the dy recombination is three TMEM column blocks of three different accumulators added in the
// epilogue registers -- no data moves between threads because an M tile is ONE h-line (128 consecutive w voxels).
//
// A CTA owns (b, d, a range of h-lines, a 128-wide w tile, a 32-channel block of Cout).  Per step h' it needs the three
// input lines (d-1, d, d+1) x h' (TMA boxes [130 w][chunk], halo columns and out-of-volume lines zero-filled = padding);
// a dx tap is a one-row shift of the descriptor start.  All weight tiles stay resident in shared memory as 9 groups
// (dz, dx) of [96 = (dy, co)][Cin] K-major.  Four Q accumulators (96 TMEM columns each) rotate: the MMAs of line h'+1
// overlap the epilogue of output line h'-1.
//
// Narrow maps (36 <= W <= 64, the second U-Net level): an M tile is TWO depth planes (d, d+1) of one 64-wide h-line,
// interleaved row by row -- the TMA box is taken over (C, D = 2, W = 66) with D as the faster shared-memory dimension
// (tensor-map dimensions may be listed in any order), so tile row = 2 * w + plane, a dx tap is a TWO-row shift of the
// descriptor start, and everything else (dz = another box one plane further, dy merged into N, Q ring, lane-local
// epilogue) is unchanged: thread m of the epilogue owns voxel (w = m / 2, plane d + m % 2).
//
// Warp roles: 0 = line producer (TMA), 1 = TMEM owner + MMA issuer, 2 = weight loader, 3.. = EW epilogue warps
// (EW / 4 warps per TMEM lane quarter, LN_CPT of the 32 output channels each).
#include "umma.cuh"

namespace mtb {

using namespace um;

// EW = epilogue warps: 8 (16 output channels per thread) or 16 (8 channels per thread; twice the warps to hide the
// TMEM-load / store latencies of the per-line epilogue).  Threads = 3 service warps + EW epilogue warps.
constexpr int ln_threads(int ew) { return 96 + 32 * ew; }
constexpr int LN_MAX_STAGES = 6;
constexpr int LN_QSLOTS = 4;
constexpr int LN_WROWS = 130;   // 128 output columns + halo
constexpr int LN_BN = 32;       // Cout block per CTA; N = 3 * LN_BN

struct LineParams {
  CUtensorMap a_map, w_map;
  void* out;
  const float* bias;
  double* stats;
  int B, D, H, W;
  int out_ldc, out_coff, Cout;
  int kcw, nchunk;
  int sub_bytes;                 // one chunk of one line, 1024-aligned
  int line_bytes;                // nchunk * sub_bytes
  int stage_bytes, stage_tx;     // ndz lines
  int stages;
  int wchunk_bytes, wgroup_bytes;  // [96][kcw] tile, one (dz,dx) group = nchunk tiles
  int ngroups;
  int ndz, dz0;                  // planes d+dz0 .. d+dz0+ndz-1 are staged
  int grp_dzslot[9], grp_dxrow[9];    // which staged line, row offset (dx + 1)
  int grp_widx[9][3];            // weight slice of (group, dy = j - 1)
  int nhr, hlen, ntw;
  int P, wt, dgroups;            // planes per M tile (1 or 2), w voxels per tile (128 / P), D / P
  int units;                     // work units walked by the persistent CTAs
  int accumulate, is_f16;
  int in_split;                  // planar input halves: chunk c is sample b + c * B of a [2B] tensor (channel 0)
  int out_split;                 // planar output halves: Cout block n0 goes to half n0 / out_split
  int dbg;                       // experiments only (MTB200_LINE_DBG): bit 0 = read one accumulator block per line
  // fused InstanceNorm-backward reduction (RED kernels, see mtb200_conv_params::red)
  const void* red_y;
  const float4* red_xform;
  const float2* red_meanrstd;
  double* red;
  int red_ldc, red_coff;
};

__device__ __forceinline__ uint64_t ln_desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | (uint64_t)lo; }
__device__ __forceinline__ uint32_t ln_kmajor_hi(uint32_t row_bytes, uint32_t sbo) {
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
}

// RED = data-gradient launch with the InstanceNorm-backward reduction of the PRODUCING layer fused into the epilogue: the
// thread that holds g = d(loss)/d(activation) of a voxel also reads the 16 raw conv outputs y of that voxel and accumulates
// sum dv, sum dv * xhat per (b, channel) exactly like the forward statistics (fixed order inside the CTA, fp64 atomics
// across CTAs) -- the separate pass over g and y (mtb200_in_bwd_reduce, 4 B per element) disappears.
// NCH = channel chunks per line (compile time: the MMA issue loop must stay a straight run of instructions): 1, or 2 for
// planar input halves.
template <typename T, int ROWB, int EW, bool RED, int NCH>
__global__ void __launch_bounds__(ln_threads(EW), 1) conv_line_umma_kernel(const __grid_constant__ LineParams p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  constexpr int LN_CPT = 32 / (EW / 4);  // output channels per epilogue thread
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t st_full[LN_MAX_STAGES], st_empty[LN_MAX_STAGES];
  __shared__ __align__(8) uint64_t q_full[LN_QSLOTS], q_empty[LN_QSLOTS];
  __shared__ __align__(8) uint64_t w_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[LN_BN];
  __shared__ float4 s_rc[LN_BN];     // RED: {scale, shift, rstd, -mean * rstd} of (current sample, channel)
  __shared__ float s_rslope[LN_BN];  // RED: LeakyReLU slope

  constexpr int KSTEPS = ROWB / 32;
  constexpr uint32_t NCOLS = 3 * LN_BN;  // accumulator width
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* st_base = dsmem;
  uint8_t* w_base = dsmem + (size_t)p.stages * p.stage_bytes;

  // PERSISTENT CTAs: blockIdx.x strides over the work units (b, d, h range, w tile); barriers, TMEM, weights and the
  // pipeline state (global step counter `gs` -> stage ring / Q ring positions and parities) live for the whole kernel,
  // so consecutive units overlap (the producer runs ahead into the next unit while the epilogue drains this one).
  const int n0 = blockIdx.y * LN_BN;
  struct Unit { int b, d, hs, he, w0, hfirst, nsteps; };
  auto decode = [&](int u) {
    Unit t;
    const int hr = u % p.nhr; u /= p.nhr;
    const int twi = u % p.ntw; u /= p.ntw;
    t.d = (u % p.dgroups) * p.P;
    t.b = u / p.dgroups;
    t.hs = hr * p.hlen;
    t.he = min(p.H, t.hs + p.hlen);
    t.w0 = twi * p.wt;
    t.hfirst = max(t.hs - 1, 0);
    t.nsteps = min(t.he, p.H - 1) - t.hfirst + 1;
    return t;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], 1); }
    for (int i = 0; i < LN_QSLOTS; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], EW); }
    mbar_init(&w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < LN_BN) s_bias[threadIdx.x] = p.bias ? p.bias[n0 + threadIdx.x] : 0.f;
  if (warp == 1) tmem_alloc(&tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== line producer =====
    uint32_t gs = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const Unit t = decode(u);
      for (int s = 0; s < t.nsteps; ++s, ++gs) {
        const uint32_t slot = gs % (uint32_t)p.stages;
        mbar_wait(&st_empty[slot], ((gs / (uint32_t)p.stages) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&st_full[slot], (uint32_t)p.stage_tx);
          uint8_t* dst = st_base + (size_t)slot * p.stage_bytes;
          for (int z = 0; z < p.ndz; ++z)
            for (int c = 0; c < p.nchunk; ++c) {
              uint8_t* sub = dst + (size_t)z * p.line_bytes + (size_t)c * p.sub_bytes;
              if (p.in_split)  // planar halves: chunk c = the whole channel range of half c (sample index b + c * B)
                tma_load_5d(sub, &p.a_map, &st_full[slot], 0, t.w0 - 1, t.hfirst + s, t.d + p.dz0 + z, t.b + c * p.B);
              else if (p.P == 1)  // map dims (C, W, H, D, B)
                tma_load_5d(sub, &p.a_map, &st_full[slot], c * p.kcw, t.w0 - 1, t.hfirst + s, t.d + p.dz0 + z, t.b);
              else           // map dims (C, D, W, H, B): planes d+dz, d+dz+1 interleaved row by row
                tma_load_5d(sub, &p.a_map, &st_full[slot], c * p.kcw, t.d + p.dz0 + z, t.w0 - 1, t.hfirst + s, t.b);
            }
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ===== weight loader: every (group, dy, chunk) tile once =====
    if (elect_one()) {
      mbar_expect_tx(&w_full, (uint32_t)(p.ngroups * 3 * p.nchunk * LN_BN * ROWB));
      for (int g = 0; g < p.ngroups; ++g)
        for (int c = 0; c < p.nchunk; ++c)
          for (int j = 0; j < 3; ++j)
            tma_load_3d(w_base + (size_t)g * p.wgroup_bytes + (size_t)c * p.wchunk_bytes + (size_t)j * (LN_BN * ROWB),
                        &p.w_map, &w_full, c * p.kcw, n0, p.grp_widx[g][j]);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = idesc_f16(p.is_f16 != 0, NCOLS, false, false);
    const uint32_t hi = ln_kmajor_hi(ROWB, 8u * ROWB);
    const uint32_t st16 = __shfl_sync(0xffffffffu, (smem_u32(st_base) & 0x3FFFFu) >> 4, 0);
    const uint32_t w16 = __shfl_sync(0xffffffffu, (smem_u32(w_base) & 0x3FFFFu) >> 4, 0);
    const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, line16 = (uint32_t)p.line_bytes >> 4;
    const uint32_t wgroup16 = (uint32_t)p.wgroup_bytes >> 4;
    const uint32_t sub16 = (uint32_t)p.sub_bytes >> 4, wchunk16 = (uint32_t)p.wchunk_bytes >> 4;
    const int ngroups = (p.dbg & 8) ? 1 : p.ngroups;
    // per-group descriptor offsets live in (uniform) registers: the issue loop is a handful of adds per MMA.  The single
    // issuing warp pays the full latency of every dependent instruction, so nothing else may sit between two MMAs.
    uint32_t a_goff[9], b_goff[9];
#pragma unroll
    for (int g = 0; g < 9; ++g) {
      a_goff[g] = (uint32_t)p.grp_dzslot[g] * line16 + (uint32_t)(p.grp_dxrow[g] * p.P * (ROWB / 16));
      b_goff[g] = w16 + (uint32_t)g * wgroup16;
    }
    mbar_wait(&w_full, 0);
    tc_fence_after();
    uint32_t gs = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const int nsteps = decode(u).nsteps;
      for (int s = 0; s < nsteps; ++s, ++gs) {
        const uint32_t slot = gs % (uint32_t)p.stages;
        const uint32_t qs = gs % LN_QSLOTS;
        mbar_wait(&q_empty[qs], ((gs / LN_QSLOTS) & 1u) ^ 1u);
        mbar_wait(&st_full[slot], (gs / (uint32_t)p.stages) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_s = st16 + slot * stage16;
          const uint32_t dq = tmem_u + qs * NCOLS;
#pragma unroll
          for (int g = 0; g < 9; ++g) {
            if (g < ngroups) {
#pragma unroll
              for (int c = 0; c < NCH; ++c) {  // channel chunks (2 for planar halves, else 1)
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k)
                  umma_f16(dq, ln_desc64(hi, a_s + a_goff[g] + (uint32_t)c * sub16 + (uint32_t)(k * 2)),
                           ln_desc64(hi, b_goff[g] + (uint32_t)c * wchunk16 + (uint32_t)(k * 2)), idesc, (g | c | k) ? 1u : 0u);
              }
            }
          }
          umma_commit(&st_empty[slot]);
          umma_commit(&q_full[qs]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue warps 3..: TMEM lane quarter = warp % 4, channel part = (warp - 3) / 4; thread = one w column =====
    const int q = warp & 3;
    const int part = (warp - 3) >> 2;
    T* out = reinterpret_cast<T*>(p.out);
    int n0o = n0;  // channel offset of this CTA's block inside its output tensor
    if (p.out_split) {  // planar output halves: [2][B][D][H][W][out_ldc]
      out += (long long)(n0 / p.out_split) * p.B * p.D * p.H * p.W * p.out_ldc;
      n0o = n0 % p.out_split;
    }
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * LN_CPT);
    float bias[LN_CPT];  // RED (data gradient: no bias) keeps the InstanceNorm scale of its channels here instead
#pragma unroll
    for (int j = 0; j < LN_CPT; ++j) bias[j] = RED ? 0.f : s_bias[part * LN_CPT + j];
    float rshift[RED ? LN_CPT : 1];
    float rslope = 0.f;
    const bool want_stats = !RED && p.stats != nullptr;
    int cur_b = -1;
    uint32_t gs0 = 0;  // global step index of the current unit's first line
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const Unit t = decode(u);
      if (RED && t.b != cur_b) {  // per-(sample, channel) constants of the fused reduction (uniform over the epilogue warps)
        asm volatile("bar.sync 1, %0;" ::"r"(EW * 32) : "memory");
        const int et = (int)threadIdx.x - 96;
        if (et < LN_BN) {
          const long long i = (long long)t.b * p.Cout + n0 + et;
          const float4 f = p.red_xform[i];
          const float2 mr = p.red_meanrstd[i];
          s_rc[et] = make_float4(f.x, f.y, mr.y, -mr.x * mr.y);
          s_rslope[et] = f.z;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(EW * 32) : "memory");
        cur_b = t.b;
#pragma unroll
        for (int j = 0; j < LN_CPT; ++j) {  // per-channel constants live in registers for the whole sample
          const float4 c = s_rc[part * LN_CPT + j];
          bias[j] = c.x;
          rshift[j] = c.y;
        }
        rslope = s_rslope[part * LN_CPT];  // one LeakyReLU slope per layer
      }
      const int m = q * 32 + lane;  // tile row = TMEM lane
      const int b = t.b, d = t.d + (p.P == 2 ? (m & 1) : 0), hs = t.hs, he = t.he, hfirst = t.hfirst;
      const int ww = t.w0 + (p.P == 2 ? (m >> 1) : m);
      const bool wvalid = ww < p.W;
      float csum[LN_CPT], csq[LN_CPT];
#pragma unroll
      for (int j = 0; j < LN_CPT; ++j) { csum[j] = 0.f; csq[j] = 0.f; }

      // RED: the raw outputs y of line h were requested one line earlier (their global-load latency would otherwise sit
      // in the critical path of every line); emit(h) consumes them and requests line h + 1
      Raw8<T> ynext[LN_CPT / 8];
      auto ylines = [&](int h) {
        const T* yrow = reinterpret_cast<const T*>(p.red_y) +
                        ((((long long)b * p.D + d) * p.H + h) * p.W + ww) * p.red_ldc + p.red_coff + n0 + part * LN_CPT;
#pragma unroll
        for (int c8 = 0; c8 < LN_CPT / 8; ++c8) ynext[c8].load(yrow + c8 * 8);
      };
      if (RED && wvalid) ylines(hs);

      auto emit = [&](int h) {
        uint32_t r[3][LN_CPT];
        Raw8<T> yraw[LN_CPT / 8];
        if (RED && wvalid) {
#pragma unroll
          for (int c8 = 0; c8 < LN_CPT / 8; ++c8) yraw[c8] = ynext[c8];
          if (h + 1 < he) ylines(h + 1);
        }
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const int hq = h + dy;
          if (hq >= 0 && hq < p.H && !((p.dbg & 1) && dy != 0)) {
            tmem_ld_async(tlane + ((gs0 + (uint32_t)(hq - hfirst)) % LN_QSLOTS) * NCOLS + (uint32_t)((dy + 1) * LN_BN), r[dy + 1]);
          } else {  // out-of-volume line: zero contribution
#pragma unroll
            for (int j = 0; j < LN_CPT; ++j) r[dy + 1][j] = 0u;
          }
        }
        tmem_ld_fence(r[0]);
        tmem_ld_fence(r[1]);
        tmem_ld_fence(r[2]);
        // the three accumulators are in registers: Q[h-1] is not needed by any later output line
        tc_fence_before();
        __syncwarp();
        if (h - 1 >= hfirst && lane == 0) mbar_arrive(&q_empty[(gs0 + (uint32_t)(h - 1 - hfirst)) % LN_QSLOTS]);
        if (wvalid) {
          float v[LN_CPT];
#pragma unroll
          for (int j = 0; j < LN_CPT; ++j)
            v[j] = (RED ? 0.f : bias[j]) + __uint_as_float(r[0][j]) + __uint_as_float(r[1][j]) + __uint_as_float(r[2][j]);
          T* orow = out + ((((long long)b * p.D + d) * p.H + h) * p.W + ww) * p.out_ldc + p.out_coff + n0o + part * LN_CPT;
#pragma unroll
          for (int c8 = 0; c8 < LN_CPT; c8 += 8) {
            float o8[8];
            if (p.accumulate) {
              load8<T>(orow + c8, o8);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[c8 + j] += o8[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) o8[j] = v[c8 + j];
            if (!(p.dbg & 2)) store8<T>(orow + c8, o8);
          }
          if (want_stats && !(p.dbg & 4)) {
#pragma unroll
            for (int j = 0; j < LN_CPT; ++j) {
              const float x = Traits<T>::round(v[j]);
              csum[j] += x;
              csq[j] = fmaf(x, x, csq[j]);
            }
          }
          if (RED) {  // csum = sum dv, csq = sum dv * y; xhat = y * rstd - mean * rstd is applied once per unit at the flush
#pragma unroll
            for (int j = 0; j < LN_CPT; ++j) {
              const float g = Traits<T>::round(v[j]);  // what in_bwd_apply will read back
              const float yv = yraw[j >> 3].get(j & 7);
              const float dv = fmaf(yv, bias[j], rshift[j]) > 0.f ? g : g * rslope;
              csum[j] += dv;
              csq[j] = fmaf(dv, yv, csq[j]);
            }
          }
        }
      };

      for (int s = 0; s < t.nsteps; ++s) {
        const int hp = hfirst + s;
        const uint32_t g = gs0 + (uint32_t)s;
        mbar_wait(&q_full[g % LN_QSLOTS], (g / LN_QSLOTS) & 1u);
        tc_fence_after();
        if (hp - 1 >= hs) emit(hp - 1);
        if (hp == p.H - 1 && hp < he) emit(hp);  // last line of the volume: Q[H] does not exist
      }
      // emit(h) released Q[h-1] for h in [hs, he): the accumulators of lines he-1 .. hfirst+nsteps-1 are still held.
      // Hand them back (every TMEM read of this warp has completed) so the next unit's MMAs can reuse the slots.
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        for (int hq = max(he - 1, hfirst); hq < hfirst + t.nsteps; ++hq) mbar_arrive(&q_empty[(gs0 + (uint32_t)(hq - hfirst)) % LN_QSLOTS]);
      if (want_stats || RED) {  // per-(b, channel) sums of this unit -> fp64 atomics (warp-level column sums first)
        warp_colsum(csum, lane);
        warp_colsum(csq, lane);
        if (colsum_writer<LN_CPT>(lane)) {
          const int col = part * LN_CPT + colsum_column<LN_CPT>(lane);
          double* st = (RED ? p.red : p.stats) + ((long long)b * p.Cout + n0 + col) * 2;
          if (RED) {  // sum dv * xhat = rstd * sum dv * y - mean * rstd * sum dv
            const float4 c = s_rc[col];
            atomicAdd(st, (double)csum[0]);
            atomicAdd(st + 1, (double)c.z * (double)csq[0] + (double)c.w * (double)csum[0]);
          } else {
            atomicAdd(st, (double)csum[0]);
            atomicAdd(st + 1, (double)csq[0]);
          }
        }
      }
      gs0 += (uint32_t)t.nsteps;
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

template <typename T, int ROWB, int EW, bool RED, int NCH>
static cudaError_t launch_line_red(const LineParams& q, dim3 grid, int smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(conv_line_umma_kernel<T, ROWB, EW, RED, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) launch_pdl(conv_line_umma_kernel<T, ROWB, EW, RED, NCH>, dim3(grid), dim3(ln_threads(EW)), (size_t)(smem), s, q);
  return e;
}
template <typename T, int ROWB, int EW>
static cudaError_t launch_line_ew(const LineParams& q, dim3 grid, int smem, cudaStream_t s) {
  if (q.nchunk == 2) {  // planar input halves (64-byte rows, forward launches)
    if (ROWB != 64 || q.red) return cudaErrorInvalidValue;
    return launch_line_red<T, 64, EW, false, 2>(q, grid, smem, s);
  }
  return q.red ? launch_line_red<T, ROWB, EW, true, 1>(q, grid, smem, s) : launch_line_red<T, ROWB, EW, false, 1>(q, grid, smem, s);
}

// MTB200_LINE_EPI_WARPS=8|16 overrides the epilogue width (experiments); default 8 (measured faster, profiles/r1i_line_epi_ab.txt)
static int line_epi_warps() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("MTB200_LINE_EPI_WARPS");
    v = (e && atoi(e) == 16) ? 16 : 8;
  }
  return v;
}

template <typename T, int ROWB>
static cudaError_t launch_line(const LineParams& q, dim3 grid, int smem, cudaStream_t s) {
  return line_epi_warps() == 8 ? launch_line_ew<T, ROWB, 8>(q, grid, smem, s) : launch_line_ew<T, ROWB, 16>(q, grid, smem, s);
}

static inline int ln_align1k(long long v) { return (int)(((v + 1023) / 1024) * 1024); }

static int conv_line_launch(const mtb200_conv_params& p, cudaStream_t s, int w_pitch, int w_coff);

// Returns MTB200_ERR_UNSUPPORTED when the problem is outside this kernel's envelope (caller falls back).
//
// 128 input channels (the first decoder convolution of the second level, 128 -> 64 at 96x80x64: its 9 x 96 x 128 weight
// tiles per CTA would need 221 KB): TWO passes over one 64-channel half of the input each, conv(x) = conv_lo(x[:64]) +
// conv_hi(x[64:]) -- the first stores bias + partial sum in the output's 16-bit type, the second adds its half on top
// (`accumulate`) and takes the InstanceNorm statistics of the final values.  One extra rounding of the partial sum
// (<= 0.5 ulp of the 16-bit type), 0.25 GB of extra traffic.  Measured 0.91 ms against 1.01 ms on the CTA-pair per-tap
// kernel (the 64-channel halves of 256-byte rows stream slower than a dense 64-channel tensor): not worth a second
// rounding, so OFF unless MTB200_LINE_2PASS=1.
int conv_line_umma(const mtb200_conv_params& p, cudaStream_t s) {
  static const int two_pass = [] { const char* e = getenv("MTB200_LINE_2PASS"); return (e && atoi(e) == 1) ? 1 : 0; }();
  if (two_pass && p.Cin == 128 && !p.in_split && !p.red && !p.xform) {
    mtb200_conv_params a = p;
    a.Cin = 64; a.stats = nullptr;
    const int r = conv_line_launch(a, s, 128, 0);
    if (r != MTB200_OK) return r;  // outside the envelope: nothing was launched
    mtb200_conv_params b = p;
    b.Cin = 64; b.in_coff = p.in_coff + 64; b.bias = nullptr; b.accumulate = 1;
    const int r2 = conv_line_launch(b, s, 128, 64);
    if (r2 == MTB200_ERR_UNSUPPORTED) { set_error("conv_line: second half of a two-pass launch refused"); return MTB200_ERR_CUDA; }
    return r2;
  }
  return conv_line_launch(p, s, 0, 0);
}

// w_pitch / w_coff: the weights are a [tap][Cout][w_pitch] array of which this launch uses channels [w_coff, w_coff + Cin)
// (0 = dense [tap][Cout][Cin])
static int conv_line_launch(const mtb200_conv_params& p, cudaStream_t s, int w_pitch, int w_coff) {
  if (p.ngroups != 1) return MTB200_ERR_UNSUPPORTED;
  for (int k = 0; k < 3; ++k)
    if (p.is[k] != 1 || p.os[k] != 1 || p.group_ooff[0][k] != 0) return MTB200_ERR_UNSUPPORTED;
  if (p.Do != p.Di || p.Ho != p.Hi || p.Wo != p.Wi || p.Dof != p.Do || p.Hof != p.Ho || p.Wof != p.Wo)
    return MTB200_ERR_UNSUPPORTED;
  if (p.Cin != 16 && p.Cin != 32 && p.Cin != 64) return MTB200_ERR_UNSUPPORTED;  // one chunk, or two planar halves
  if (p.in_split && (p.in_split != 32 || p.Cin != 64 || p.in_coff != 0 || p.Wo < 72 || p.red)) return MTB200_ERR_UNSUPPORTED;
  if (p.out_split && (p.out_split % LN_BN || p.Cout != 2 * p.out_split || p.out_coff != 0 || p.red)) return MTB200_ERR_UNSUPPORTED;
  if (p.Cout % LN_BN) return MTB200_ERR_UNSUPPORTED;
  // an M tile is one h-line of 128 w voxels, or (narrow maps) two depth planes of one 64-wide h-line
  static int pair_ok = -1;  // MTB200_LINE_PAIR=0 disables the two-plane variant; cleared if its tensor map is refused
  if (pair_ok < 0) { const char* e = getenv("MTB200_LINE_PAIR"); pair_ok = (e && atoi(e) == 0) ? 0 : 1; }
  const int P = p.Wo >= 72 ? 1 : 2;
  if (p.Ho < 4) return MTB200_ERR_UNSUPPORTED;
  if (P == 2 && (!pair_ok || p.Wo < 36 || p.Wo > 64 || (p.Do & 1) || p.Cin < 32)) return MTB200_ERR_UNSUPPORTED;

  static thread_local LineParams q;
  memset(&q, 0, sizeof(q));
  // groups (dz, dx), each with all three dy taps
  int gid[3][3];
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c) gid[a][c] = -1;
  int dzmin = 1, dzmax = -1;
  for (int t = 0; t < p.ntaps; ++t) {
    for (int k = 0; k < 3; ++k)
      if (p.tap_off[t][k] < -1 || p.tap_off[t][k] > 1) return MTB200_ERR_UNSUPPORTED;
    dzmin = min(dzmin, p.tap_off[t][0]); dzmax = max(dzmax, p.tap_off[t][0]);
  }
  q.dz0 = dzmin; q.ndz = dzmax - dzmin + 1;
  for (int g = 0; g < 9; ++g)
    for (int j = 0; j < 3; ++j) q.grp_widx[g][j] = -1;
  for (int t = 0; t < p.ntaps; ++t) {
    const int dz = p.tap_off[t][0], dy = p.tap_off[t][1], dx = p.tap_off[t][2];
    int& g = gid[dz + 1][dx + 1];
    if (g < 0) {
      g = q.ngroups++;
      q.grp_dzslot[g] = dz - dzmin;
      q.grp_dxrow[g] = dx + 1;
    }
    if (q.grp_widx[g][dy + 1] >= 0) return MTB200_ERR_UNSUPPORTED;  // duplicate tap
    q.grp_widx[g][dy + 1] = p.tap_widx[t];
  }
  for (int g = 0; g < q.ngroups; ++g)
    for (int j = 0; j < 3; ++j)
      if (q.grp_widx[g][j] < 0) return MTB200_ERR_UNSUPPORTED;  // needs all three dy taps of every (dz, dx)

  q.kcw = p.in_split ? p.in_split : (p.Cin < 64 ? p.Cin : 64);
  q.nchunk = p.Cin / q.kcw;
  q.in_split = p.in_split; q.out_split = p.out_split;
  const int rowb = q.kcw * 2;
  q.P = P;
  q.wt = 128 / P;
  q.dgroups = p.Do / P;
  const int box_rows = (q.wt + 2) * P;  // 130, or 66 w positions x 2 planes = 132
  q.sub_bytes = ln_align1k((long long)box_rows * rowb);
  q.line_bytes = q.sub_bytes * q.nchunk;
  q.stage_bytes = q.line_bytes * q.ndz;
  q.stage_tx = q.ndz * q.nchunk * box_rows * rowb;
  q.wchunk_bytes = ln_align1k(3LL * LN_BN * rowb);
  q.wgroup_bytes = q.wchunk_bytes * q.nchunk;
  const int wbytes = q.wgroup_bytes * q.ngroups;
  const int smem_budget = 224 * 1024;
  q.stages = min(LN_MAX_STAGES, (smem_budget - wbytes) / q.stage_bytes);
  if (q.stages < 2) return MTB200_ERR_UNSUPPORTED;

  if (P == 1) {
    // planar halves: one tensor of 2B samples with in_split channels each
    cuuint64_t dims[5] = {(cuuint64_t)(p.in_split ? p.in_split : p.Cin), (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.Di,
                          (cuuint64_t)(p.in_split ? 2 * p.B : p.B)};
    cuuint64_t strides[4] = {(cuuint64_t)p.in_ldc * 2, (cuuint64_t)p.Wi * p.in_ldc * 2,
                             (cuuint64_t)p.Hi * p.Wi * p.in_ldc * 2, (cuuint64_t)p.Di * p.Hi * p.Wi * p.in_ldc * 2};
    cuuint32_t box[5] = {(cuuint32_t)q.kcw, (cuuint32_t)LN_WROWS, 1, 1, 1};
    if (!umma_encode_map(&q.a_map, p.dtype, 5, (uint8_t*)p.in + (size_t)p.in_coff * 2, dims, strides, box, rowb))
      return MTB200_ERR_CUDA;
  } else {
    // dimensions listed as (C, D, W, H, B): the box [kcw][2 planes][wt + 2] lands in shared memory with the plane index
    // varying faster than w, i.e. tile row = 2 * w + plane (strides need not be sorted; out-of-volume = zero fill)
    cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Di, (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.B};
    cuuint64_t strides[4] = {(cuuint64_t)p.Hi * p.Wi * p.in_ldc * 2, (cuuint64_t)p.in_ldc * 2,
                             (cuuint64_t)p.Wi * p.in_ldc * 2, (cuuint64_t)p.Di * p.Hi * p.Wi * p.in_ldc * 2};
    cuuint32_t box[5] = {(cuuint32_t)q.kcw, 2, (cuuint32_t)(q.wt + 2), 1, 1};
    if (!umma_encode_map(&q.a_map, p.dtype, 5, (uint8_t*)p.in + (size_t)p.in_coff * 2, dims, strides, box, rowb)) {
      pair_ok = 0;  // this driver insists on sorted strides: use the other kernels from now on
      return MTB200_ERR_UNSUPPORTED;
    }
  }
  {
    int n_widx = 0;
    for (int t = 0; t < p.ntaps; ++t) n_widx = max(n_widx, p.tap_widx[t] + 1);
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, (cuuint64_t)n_widx};
    const long long wp = w_pitch ? w_pitch : p.Cin;
    cuuint64_t strides[2] = {(cuuint64_t)wp * 2, (cuuint64_t)wp * p.Cout * 2};
    cuuint32_t box[3] = {(cuuint32_t)q.kcw, (cuuint32_t)LN_BN, 1};
    if (!umma_encode_map(&q.w_map, p.dtype, 3, (uint8_t*)p.w + (size_t)w_coff * 2, dims, strides, box, rowb)) return MTB200_ERR_CUDA;
  }
  q.out = p.out; q.bias = p.bias; q.stats = p.stats;
  // fused reduction of the producing layer's InstanceNorm backward (data-gradient launches; MTB200_FUSE_RED=0: off)
  static int fuse_red = -1;
  if (fuse_red < 0) { const char* e = getenv("MTB200_FUSE_RED"); fuse_red = (e && atoi(e) == 0) ? 0 : 1; }
  if (p.red && fuse_red && !p.stats && p.red_y && p.red_xform && p.red_meanrstd && p.red_ldc % 8 == 0 && p.red_coff % 8 == 0) {
    q.red = p.red; q.red_y = p.red_y;
    q.red_xform = reinterpret_cast<const float4*>(p.red_xform);
    q.red_meanrstd = reinterpret_cast<const float2*>(p.red_meanrstd);
    q.red_ldc = p.red_ldc; q.red_coff = p.red_coff;
  }
  q.B = p.B; q.D = p.Do; q.H = p.Ho; q.W = p.Wo;
  q.out_ldc = p.out_ldc; q.out_coff = p.out_coff; q.Cout = p.Cout;
  q.accumulate = p.accumulate;
  q.is_f16 = p.dtype == MTB200_F16;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MTB200_LINE_DBG"); dbg = e ? atoi(e) : 0; } q.dbg = dbg; }
  q.ntw = (p.Wo + q.wt - 1) / q.wt;
  const int ny = p.Cout / LN_BN;
  // split H into ranges so that the units divide evenly over the persistent CTAs (one per SM and Cout block); every
  // range recomputes two halo lines
  const int gx_max = max(1, num_sms() / ny);
  {
    const long long base = (long long)p.B * q.dgroups * q.ntw;
    double best = -1;
    int best_nhr = 1;
    for (int nhr = 1; nhr <= max(1, p.Ho / 8); ++nhr) {
      const int hlen = (p.Ho + nhr - 1) / nhr;
      if ((p.Ho + hlen - 1) / hlen != nhr) continue;
      const long long units = base * nhr;
      const long long gx = units < gx_max ? units : (long long)gx_max;
      const long long rounds = (units + gx - 1) / gx;
      const double eff = (double)units / (double)(rounds * gx) * hlen / (hlen + 2.0);
      if (eff > best) { best = eff; best_nhr = nhr; }
    }
    q.nhr = best_nhr;
    q.hlen = (p.Ho + q.nhr - 1) / q.nhr;
  }
  const long long units = (long long)p.B * q.dgroups * q.ntw * q.nhr;
  MTB_REQUIRE(units < (1LL << 31), "conv_line: too many work units");
  q.units = (int)units;
  const int smem = max(116 * 1024, q.stages * q.stage_bytes + wbytes + 1024);  // one CTA per SM (512 TMEM columns each)
  dim3 grid((unsigned)(units < gx_max ? units : (long long)gx_max), ny, 1);
  cudaError_t e;
  if (p.dtype == MTB200_BF16) {
    e = rowb == 128 ? launch_line<__nv_bfloat16, 128>(q, grid, smem, s)
                    : (rowb == 64 ? launch_line<__nv_bfloat16, 64>(q, grid, smem, s)
                                  : launch_line<__nv_bfloat16, 32>(q, grid, smem, s));
  } else {
    e = rowb == 128 ? launch_line<__half, 128>(q, grid, smem, s)
                    : (rowb == 64 ? launch_line<__half, 64>(q, grid, smem, s) : launch_line<__half, 32>(q, grid, smem, s));
  }
  if (e != cudaSuccess) { set_error("conv_line: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  return check_launch(q.red ? "conv_line_umma+red" : "conv_line_umma");
}

}  // namespace mtb
