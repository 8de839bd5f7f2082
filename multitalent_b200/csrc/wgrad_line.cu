// Line-streaming tcgen05 weight gradient for the narrow full-resolution 3x3x3 layers (stride 1, taps in [-1,1]^3,
// Cout tile 32, Cin chunks of <= 32 channels) -- the autograd of conv_line.cu's layers w.r.t. their weights:
//     dW[(dz,dy,dx)][co][ci] = sum_{b,d,h,w} dY[b,d,h,w][co] * X[b,d+dz,h+dy,w+dx][ci]
//
// GEMM view with K = voxels along w (16 per MMA), both operands MN-major straight out of the NDHWC tensors:
//   A = an X line [K = w][M = 128]: the M blocks are ROW-SHIFTED VIEWS of the same shared-memory line (LBO = one row),
//       block j = the line shifted by j voxels = tap dx = j - 1  (4 blocks of 32 channels, or 8 of 16; dx <= 1 are used)
//   B = three consecutive dY lines [K = w][N = 96]: block jj = line h'-1+jj = tap dy = 1 - jj  (LBO = one line)
//   D_dz [128 = (dx, ci)][96 = (dy, co)] fp32 in TMEM, one accumulator per dz, resident for the whole kernel.
// tools/umma_probe_mn.cu verified on the B200 that MN-major descriptors accept any start row, any 8-row-group stride and
// an M-block stride of one row (overlapping blocks).  One MMA therefore does 9 taps' worth of useful work (of 12).
//
// Persistent CTAs (one per SM) walk a list of (b, d, h-range, w-tile) units; every step h' stages the three X lines
// (d-1, d, d+1) x h' and the dY lines h'-1 .. h'+1 of plane d (TMA, out-of-volume = zero fill), so steps are independent
// and the pipeline never drains between units.  Epilogue once per CTA: fp32 atomics into dW.
//
// The same kernel covers the 1x1x1 heads (generic_UNet.py:349-351): window of ONE dY line (nb = 1, only the tap
// (0,0,0) is written back) and a 48-channel Cout as two column segments (32 channels / SWIZZLE_64B + 16 / SWIZZLE_32B,
// two MMAs per K step into adjacent TMEM columns) -- that layer is pure HBM streaming.
//
// Warp roles (6 warps): 0 = producer (TMA), 1 = TMEM owner + MMA issuer, 2..5 = epilogue.
#include "umma.cuh"

namespace mtb {

using namespace um;

constexpr int WL_THREADS = 192;
constexpr int WL_MAX_STAGES = 6;
constexpr int WL_BN = 32;

struct WgradLineParams {
  CUtensorMap x_map, dy_map[2];
  int nb;                    // dY lines in the window: 3 (dy = +1, 0, -1) or 1 (all taps have dy == 0)
  int nseg;                  // Cout segments of this CTA: {32} or {32, 16}
  int seg_w[2], seg_c0[2], seg_off[2];  // channels, first channel, byte offset inside the dY window
  int ntot;                  // TMEM columns per dz accumulator = nb * sum(seg_w)
  int cout_blk;              // channels covered by one CTA (grid.z stride)
  float* dw;
  int B, D, H, W;
  int Cin, Cout;             // padded channel counts (dw strides)
  int kcw;                   // channels of this launch's X chunk (16 or 32)
  int xline_bytes;           // one staged X line, 1024-aligned
  int ybase;                 // offset of the dY window inside a stage
  int stage_bytes, stage_tx, stages;
  int dbg, x_tx;             // experiments only (MTB200_WLINE_DBG): bit 0 = skip the dY loads, bit 1 = one K step per line
  int ndz, dz0;
  int nhr, hlen, ntw;
  int wt, nkk;               // w tile (128 / 64 / 32 voxels of one line) and its number of 16-voxel K steps
  long long units;
  int npy, npz, nslots;      // (Cin chunk, Cout block) pairs and the number of CTA slots that walk the units
  int cosched;               // 1: the pairs of a unit run side by side (default); 0: pair-major grid (A/B measurements)
  int in_split;              // planar X halves: Cin chunk i is sample b + i * B of a [2B] tensor (channel 0)
  int lut[27];               // [dz+1][dy+1][dx+1] -> weight slice or -1
  int is_f16;
};

__device__ __forceinline__ uint64_t wl_desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | (uint64_t)lo; }

// ROWB = bytes per X row = kcw * 2 (64: four 32-channel blocks, 32: eight 16-channel blocks); NSEG = Cout segments per CTA
template <int ROWB, int NSEG>
__global__ void __launch_bounds__(WL_THREADS, 1) wgrad_line_umma_kernel(const __grid_constant__ WgradLineParams p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t st_full[WL_MAX_STAGES], st_empty[WL_MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t tmem_slot;

  const uint32_t NCOLS = (uint32_t)p.ntot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  // CTA -> (pair, slot).  The pairs of ONE unit sit in neighbouring CTAs and run at the same time, so the X chunk / dY
  // block a unit needs is fetched from DRAM once and served to the other pairs by the L2; pairs along grid.y / grid.z ran
  // one after the other and re-streamed both tensors from DRAM per pair (r2j ncu: 4.5 GB for 0.76 GB of operands).
  const int npairs = p.npy * p.npz;
  const int pair = p.cosched ? (int)blockIdx.x % npairs : (int)blockIdx.x / p.nslots;
  const int slot = p.cosched ? (int)blockIdx.x / npairs : (int)blockIdx.x % p.nslots;
  const int c0 = (pair % p.npy) * p.kcw;
  const int n0 = (pair / p.npy) * p.cout_blk;
  const int xc0 = p.in_split ? 0 : c0, xb0 = p.in_split ? (pair % p.npy) * p.B : 0;

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

  if (warp == 0) {
    // ===== producer =====
    uint32_t sc = 0;  // global step counter of this CTA
    for (long long u = slot; u < p.units; u += p.nslots) {
      long long t = u;
      const int hr = (int)(t % p.nhr); t /= p.nhr;
      const int twi = (int)(t % p.ntw); t /= p.ntw;
      const int d = (int)(t % p.D);
      const int b = (int)(t / p.D);
      const int hs = hr * p.hlen, he = min(p.H, hs + p.hlen);
      const int w0 = twi * p.wt;
      for (int hp = hs; hp < he; ++hp, ++sc) {
        const uint32_t slot = sc % (uint32_t)p.stages;
        mbar_wait(&st_empty[slot], ((sc / (uint32_t)p.stages) & 1u) ^ 1u);
        if (elect_one()) {
          uint8_t* dst = dsmem + (size_t)slot * p.stage_bytes;
          mbar_expect_tx(&st_full[slot], (uint32_t)((p.dbg & 1) ? p.x_tx : p.stage_tx));
          for (int z = 0; z < p.ndz; ++z)
            tma_load_5d(dst + (size_t)z * p.xline_bytes, &p.x_map, &st_full[slot], xc0, w0 - 1, hp, d + p.dz0 + z, b + xb0);
          for (int sg = 0; sg < p.nseg && !(p.dbg & 1); ++sg)  // lines hp-1, hp, hp+1 (or hp alone)
            tma_load_5d(dst + p.ybase + p.seg_off[sg], &p.dy_map[sg], &st_full[slot], n0 + p.seg_c0[sg], w0,
                        hp - (p.nb == 3 ? 1 : 0), d, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t fmt = p.is_f16 ? 0u : 1u;
    // D = f32, A/B 16-bit, both MN-major (bits 15, 16), N = nb * segment width, M = 128
    const uint32_t idesc0 = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((128u >> 4) << 24);
    const uint32_t layout_a = ROWB == 64 ? 4u : 6u;  // SWIZZLE_64B / SWIZZLE_32B
    const uint32_t hi_a = ((8u * ROWB) >> 4) | (1u << 14) | (layout_a << 29);
    const uint32_t lbo_a = ((uint32_t)ROWB >> 4) << 16;
    uint32_t idesc_s[NSEG], hi_b[NSEG], lbo_b[NSEG], yrow[NSEG], yoff16[NSEG], colo[NSEG];
#pragma unroll
    for (int sg = 0; sg < NSEG; ++sg) {
      const uint32_t w = sg == 0 ? 32u : 16u;                  // segment widths are fixed: 32 (+ 16)
      yrow[sg] = w * 2;                                        // bytes per dY row of this segment
      idesc_s[sg] = idesc0 | ((((uint32_t)p.nb * w) >> 3) << 17);
      hi_b[sg] = ((8u * yrow[sg]) >> 4) | (1u << 14) | ((w == 32 ? 4u : 6u) << 29);
      lbo_b[sg] = (((uint32_t)p.wt * yrow[sg]) >> 4) << 16;     // LBO = one dY line
      yoff16[sg] = (uint32_t)(p.ybase + p.seg_off[sg]) >> 4;
      colo[sg] = sg == 0 ? 0u : (uint32_t)(p.nb * 32);
    }
    const uint32_t s16 = __shfl_sync(0xffffffffu, (smem_u32(dsmem) & 0x3FFFFu) >> 4, 0);
    const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, xline16 = (uint32_t)p.xline_bytes >> 4;
    const int ndz = p.ndz, nkk = (p.dbg & 2) ? 1 : p.nkk;
    uint32_t sc = 0;
    for (long long u = slot; u < p.units; u += p.nslots) {
      const int hr = (int)(u % p.nhr);
      const int hs = hr * p.hlen, he = min(p.H, hs + p.hlen);
      for (int hp = hs; hp < he; ++hp, ++sc) {
        const uint32_t slot = sc % (uint32_t)p.stages;
        mbar_wait(&st_full[slot], (sc / (uint32_t)p.stages) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_s = (s16 + slot * stage16) | lbo_a;
          uint32_t b_s[NSEG];
#pragma unroll
          for (int sg = 0; sg < NSEG; ++sg) b_s[sg] = (s16 + slot * stage16 + yoff16[sg]) | lbo_b[sg];
          const uint32_t acc = sc > 0 ? 1u : 0u;
          // kk-major order: consecutive MMAs go to DIFFERENT accumulators (dz / segment), so they pipeline in the tensor
          // core instead of each waiting for the previous accumulate into the same TMEM columns
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (kk < nkk) {
#pragma unroll
              for (int z = 0; z < 3; ++z) {
                if (z < ndz) {
#pragma unroll
                  for (int sg = 0; sg < NSEG; ++sg)
                    umma_f16(tmem_u + (uint32_t)z * NCOLS + colo[sg],
                             wl_desc64(hi_a, a_s + (uint32_t)z * xline16 + (uint32_t)(kk * ROWB)),
                             wl_desc64(hi_b[sg], b_s[sg] + (uint32_t)kk * yrow[sg]), idesc_s[sg], kk ? 1u : acc);
                }
              }
            }
          }
          umma_commit(&st_empty[slot]);
        }
        __syncwarp();
      }
    }
    if (have_work) {
      if (elect_one()) umma_commit(&acc_full);
      __syncwarp();
    }
  } else if (have_work) {
    // ===== epilogue: TMEM -> fp32 atomics into dW[widx][co][ci] =====
    constexpr int CB = ROWB / 2;  // channels per M block
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int j = m / CB, ci = m % CB;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    // every lane takes part in the .sync.aligned TMEM loads; only rows of used blocks (dx = j - 1 <= 1) write back
    if ((q * 32) / CB <= 2) {
      const int nchunks = p.ntot / 16;
      const int seg0_cols = p.nb * p.seg_w[0];
      for (int z = 0; z < p.ndz; ++z) {
        const int dzi = p.dz0 + z + 1;
        for (int c16 = 0; c16 < nchunks; ++c16) {
          uint32_t r[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)z * NCOLS + (uint32_t)(c16 * 16), r);
          // column -> (segment, window line jj, channel)
          const int col = c16 * 16;
          const int sg = col < seg0_cols ? 0 : 1;
          const int rel = col - (sg ? seg0_cols : 0);
          const int jj = rel / p.seg_w[sg];
          const int co = p.seg_c0[sg] + rel % p.seg_w[sg];
          const int dyi = p.nb == 3 ? 2 - jj : 1;  // dy + 1
          const int widx = j <= 2 ? p.lut[(dzi * 3 + dyi) * 3 + j] : -1;
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
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

static inline int wl_align1k(long long v) { return (int)(((v + 1023) / 1024) * 1024); }
static int env_int_wl(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Returns MTB200_ERR_UNSUPPORTED when the problem is outside this kernel's envelope (caller falls back).
int wgrad_line_umma(const mtb200_wgrad_params& p, cudaStream_t s) {
  if (p.ngroups != 1 || p.xform) return MTB200_ERR_UNSUPPORTED;
  for (int k = 0; k < 3; ++k)
    if (p.is[k] != 1 || p.os[k] != 1 || p.group_ooff[0][k] != 0) return MTB200_ERR_UNSUPPORTED;
  if (p.Do != p.Di || p.Ho != p.Hi || p.Wo != p.Wi || p.Dof != p.Do || p.Hof != p.Ho || p.Wof != p.Wo)
    return MTB200_ERR_UNSUPPORTED;
  if (p.Cin != 16 && p.Cin % 32 != 0) return MTB200_ERR_UNSUPPORTED;
  if (p.in_split && (p.in_split != 32 || p.Cin != 64 || p.in_coff != 0)) return MTB200_ERR_UNSUPPORTED;
  // every (Cin chunk, Cout block) pair re-streams both operands from L2: worth it while the pair count is small
  const bool cout48 = p.Cout == 48 && p.Cin <= 32;  // the 47 heads on the widest maps: one column block of 32 + 16
  // MTB200_WLINE_MINW / MTB200_WLINE_MAXPAIRS: envelope knobs for A/B measurements (defaults = what measured fastest).
  // 32-wide lines (third level, 16..32 pairs) lost to the per-tap kernel while every pair re-streamed its operands from
  // DRAM (128->128 at 48x40x32: 1.25 vs 1.00 ms, r2b) and win since the pairs of a unit are co-scheduled (0.89 vs 1.04 ms,
  // r2l); 16-wide lines (64 pairs) still lose (0.92 vs 0.41 ms).
  static const int min_w = env_int_wl("MTB200_WLINE_MINW", 32), max_pairs = env_int_wl("MTB200_WLINE_MAXPAIRS", 32);
  if (!cout48 && (p.Cout % WL_BN || (p.Cin / 32) * (p.Cout / WL_BN) > max_pairs)) return MTB200_ERR_UNSUPPORTED;
  if (p.Wo < min_w || p.Wo < 16 || p.Ho < 4) return MTB200_ERR_UNSUPPORTED;
  bool all_dy0 = true;
  for (int t = 0; t < p.ntaps; ++t) all_dy0 = all_dy0 && p.tap_off[t][1] == 0;
  if (p.ntaps < 9 && !(p.ntaps == 1 && all_dy0)) return MTB200_ERR_UNSUPPORTED;

  static thread_local WgradLineParams q;
  memset(&q, 0, sizeof(q));
  for (int i = 0; i < 27; ++i) q.lut[i] = -1;
  int dzmin = 1, dzmax = -1;
  for (int t = 0; t < p.ntaps; ++t) {
    for (int k = 0; k < 3; ++k)
      if (p.tap_off[t][k] < -1 || p.tap_off[t][k] > 1) return MTB200_ERR_UNSUPPORTED;
    int& e = q.lut[((p.tap_off[t][0] + 1) * 3 + (p.tap_off[t][1] + 1)) * 3 + (p.tap_off[t][2] + 1)];
    if (e >= 0) return MTB200_ERR_UNSUPPORTED;
    e = p.tap_widx[t];
    dzmin = min(dzmin, p.tap_off[t][0]); dzmax = max(dzmax, p.tap_off[t][0]);
  }
  q.dz0 = dzmin; q.ndz = dzmax - dzmin + 1;
  q.kcw = p.Cin < 32 ? p.Cin : 32;
  const int nchunk = p.Cin / q.kcw;
  const int rowb = q.kcw * 2;
  q.wt = p.Wo > 64 ? 128 : (p.Wo > 32 ? 64 : (p.Wo > 16 ? 32 : 16));  // K = the voxels of one line, 16 per MMA
  q.nkk = q.wt / 16;
  const int xrows = q.wt + 2;
  q.xline_bytes = wl_align1k((long long)(xrows + 8) * rowb);  // + rows touched by the unused shifted blocks
  q.ybase = q.xline_bytes * q.ndz;
  q.nb = all_dy0 ? 1 : 3;
  q.nseg = cout48 ? 2 : 1;
  q.seg_w[0] = 32; q.seg_c0[0] = 0; q.seg_off[0] = 0;
  q.seg_w[1] = 16; q.seg_c0[1] = 32; q.seg_off[1] = wl_align1k((long long)q.nb * q.wt * 64);
  q.cout_blk = cout48 ? 48 : WL_BN;
  int wsum = 0, ybytes = 0;
  for (int sg = 0; sg < q.nseg; ++sg) { wsum += q.seg_w[sg]; ybytes += q.nb * q.wt * q.seg_w[sg] * 2; }
  q.ntot = q.nb * wsum;
  const int ywin_bytes = q.nseg == 2 ? q.seg_off[1] + wl_align1k((long long)q.nb * q.wt * 32) : wl_align1k((long long)q.nb * q.wt * 64);
  q.stage_bytes = wl_align1k(q.ybase + ywin_bytes);
  q.stage_tx = q.ndz * xrows * rowb + ybytes;
  q.x_tx = q.ndz * xrows * rowb;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MTB200_WLINE_DBG"); dbg = e ? atoi(e) : 0; } q.dbg = dbg; }
  q.stages = min(WL_MAX_STAGES, (224 * 1024) / q.stage_bytes);
  if (q.stages < 2) return MTB200_ERR_UNSUPPORTED;
  {
    cuuint64_t dims[5] = {(cuuint64_t)(p.in_split ? p.in_split : p.Cin), (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.Di,
                          (cuuint64_t)(p.in_split ? 2 * p.B : p.B)};
    cuuint64_t strides[4] = {(cuuint64_t)p.in_ldc * 2, (cuuint64_t)p.Wi * p.in_ldc * 2,
                             (cuuint64_t)p.Hi * p.Wi * p.in_ldc * 2, (cuuint64_t)p.Di * p.Hi * p.Wi * p.in_ldc * 2};
    cuuint32_t box[5] = {(cuuint32_t)q.kcw, (cuuint32_t)xrows, 1, 1, 1};
    if (!umma_encode_map(&q.x_map, p.dtype, 5, (uint8_t*)p.x + (size_t)p.in_coff * 2, dims, strides, box, rowb))
      return MTB200_ERR_CUDA;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.Cout, (cuuint64_t)p.Wof, (cuuint64_t)p.Hof, (cuuint64_t)p.Dof, (cuuint64_t)p.B};
    cuuint64_t strides[4] = {(cuuint64_t)p.out_ldc * 2, (cuuint64_t)p.Wof * p.out_ldc * 2,
                             (cuuint64_t)p.Hof * p.Wof * p.out_ldc * 2,
                             (cuuint64_t)p.Dof * p.Hof * p.Wof * p.out_ldc * 2};
    for (int sg = 0; sg < q.nseg; ++sg) {
      cuuint32_t box[5] = {(cuuint32_t)q.seg_w[sg], (cuuint32_t)q.wt, (cuuint32_t)q.nb, 1, 1};
      if (!umma_encode_map(&q.dy_map[sg], p.dtype, 5, (uint8_t*)p.dy + (size_t)p.out_coff * 2, dims, strides, box,
                           q.seg_w[sg] * 2))
        return MTB200_ERR_CUDA;
    }
  }
  q.dw = p.dw;
  q.B = p.B; q.D = p.Do; q.H = p.Ho; q.W = p.Wo;
  q.Cin = p.Cin; q.Cout = p.Cout;
  q.is_f16 = p.dtype == MTB200_F16;
  q.in_split = p.in_split;
  q.ntw = (p.Wo + q.wt - 1) / q.wt;
  const int sms = num_sms();
  q.npy = nchunk;
  q.npz = p.Cout / q.cout_blk;
  const int npairs = q.npy * q.npz;
  q.cosched = env_int_wl("MTB200_WLINE_COSCHED", 1) != 0;
  const int slots_max = q.cosched ? max(1, sms / npairs) : sms;
  {
    // split H into ranges so that every persistent slot gets (almost) the same number of steps
    const long long base = (long long)p.B * p.Do * q.ntw;
    double best = -1;
    int best_nhr = 1;
    for (int nhr = 1; nhr <= max(1, p.Ho / 8); ++nhr) {
      const int hlen = (p.Ho + nhr - 1) / nhr;
      if ((p.Ho + hlen - 1) / hlen != nhr) continue;
      const long long units = base * nhr;
      const long long g = units < slots_max ? units : slots_max;
      const long long per = (units + g - 1) / g;
      const double eff = (double)units / (double)(per * g);
      if (eff > best + 1e-9) { best = eff; best_nhr = nhr; }
    }
    q.nhr = best_nhr;
    q.hlen = (p.Ho + q.nhr - 1) / q.nhr;
  }
  q.units = (long long)p.B * p.Do * q.ntw * q.nhr;
  q.nslots = (int)(q.units < slots_max ? q.units : slots_max);
  const int smem = max(116 * 1024, q.stages * q.stage_bytes + 1024);  // one CTA per SM (512 TMEM columns each)
  dim3 grid((unsigned)(q.nslots * npairs), 1, 1);
  cudaError_t e = cudaSuccess;
#define WL_LAUNCH(RB, NS)                                                                                         \
  do {                                                                                                            \
    e = cudaFuncSetAttribute(wgrad_line_umma_kernel<RB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);  \
    if (e == cudaSuccess) launch_pdl(wgrad_line_umma_kernel<RB, NS>, dim3(grid), dim3(WL_THREADS), (size_t)(smem), s, q);                       \
  } while (0)
  if (rowb == 64) {
    if (q.nseg == 1) WL_LAUNCH(64, 1); else WL_LAUNCH(64, 2);
  } else {
    if (q.nseg == 1) WL_LAUNCH(32, 1); else WL_LAUNCH(32, 2);
  }
#undef WL_LAUNCH
  if (e != cudaSuccess) { set_error("wgrad_line: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  return check_launch("wgrad_line_umma");
}

}  // namespace mtb
