// Group-merged tcgen05 kernel for the stride-2 LATTICE problems at the top of the U-Net: the data gradient of the
// strided 3x3x3 convolution (generic_UNet.py:126-141 `first_stride`; 64 -> 32 channels, 96x80x64 -> 192x160x128) and
// the forward ConvTranspose3d(k == s) (generic_UNet.py:335-336).  In tap-table form both have one tap GROUP per output
// residue class r in {0,1}^3 (out voxel = 2q + r) and few distinct input offsets:
//     out[2q + r][co] (+)= sum_{taps t of group r} W[widx_t][co][:] . in[q + off_t][:]
// The per-tap kernel ran every (tap, group) as its own N = 32 MMA on its own TMA tile (27 activation loads, 27 weight
// loads and 108 MMAs per 128 voxels, direct 16-byte stores).  Here
//   * the 8 groups sit SIDE BY SIDE in one 256-column TMEM accumulator, so a tap offset shared by several groups is ONE
//     activation load and one MMA per run of adjacent groups (strided dgrad: 8 loads / 14 MMA runs instead of 27 / 27;
//     transposed conv: 1 load, one N = 256 MMA);
//   * every weight tile stays resident in shared memory (27 x 4 KB), stacked so that a run's tiles form one K-major
//     [N][K] operand;
//   * the epilogue stages the (2bd x 2bh x 2bw) output brick in swizzled shared memory, one (d-parity, h-parity)
//     quarter at a time (bd x bh x 2bw voxels = 16 KB at 32 channels), and writes it with ONE 5-D TMA store per quarter
//     (reduce-add when the gradient buffer already holds the decoder's contribution), instead of half-filled sectors
//     from per-lane stores.
// Persistent CTAs, one per SM (512 TMEM columns = two accumulators).  Warp roles: 0 = TMA producer, 1 = TMEM owner + MMA
// issuer, 2..5 = epilogue of the d-parity-0 half, 6..9 = epilogue of the d-parity-1 half (own staging buffer, own bulk
// store group: the kernel is bound by this TMEM -> bf16 -> shared -> TMA-store chain, so the two halves run side by side).
#include "umma.cuh"

namespace mtb {

using namespace um;

constexpr int GM_THREADS = 320;
constexpr int GM_MAX_STAGES = 6;
constexpr int GM_MAX_OPS = 32;

struct GmParams {
  CUtensorMap a_map, w_map, o_map[4];   // o_map[2 * r0 + r1]: the output sub-lattice d = os0*q + r0, h = os1*q + r1
  int B, tiles_d, tiles_h, tiles_w, bd, bh, bw;
  uint32_t ntiles;
  int KC, Cout, ngroups, ncols;   // ncols = ngroups * Cout (accumulator width)
  int os[3];
  int grp_r[MTB200_MAX_GROUPS][3];
  int nloads;                     // distinct input offsets
  int load_off[MTB200_MAX_GROUPS * 4][3];
  int op_begin[MTB200_MAX_GROUPS * 4 + 1];  // MMA runs of load a: [op_begin[a], op_begin[a+1])
  int op_col[GM_MAX_OPS], op_n[GM_MAX_OPS], op_w16[GM_MAX_OPS], op_first[GM_MAX_OPS];
  int nwt;                        // weight tiles
  int wt_widx[MTB200_MAX_TAPS], wt_off[MTB200_MAX_TAPS];
  int stages, a_stage_bytes, w_bytes, out_half_bytes, out_mask;
  int hsplit;                     // 1: the staging buffer holds one (d, h) parity QUARTER per store (when the resident
                                  // weights leave no room for two half bricks), 0: one d-parity half
  int accumulate, is_f16;
};

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3,
                                                  int c4) {
  asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

template <typename T>
__global__ void __launch_bounds__(GM_THREADS, 1) conv_gm_umma_kernel(const __grid_constant__ GmParams p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t full_bar[GM_MAX_STAGES], empty_bar[GM_MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2], w_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_base = dsmem;
  uint8_t* w_base = a_base + (size_t)p.stages * p.a_stage_bytes;
  uint8_t* o_base = w_base + p.w_bytes;
  const uint32_t row_bytes = p.KC * 2;
  const uint32_t wtile_bytes = (uint32_t)p.Cout * row_bytes;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * p.os[0]); }
    mbar_init(&w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: all weight tiles once, then `nloads` activation bricks per tile =====
    if (elect_one()) {
      mbar_expect_tx(&w_full, (uint32_t)p.nwt * wtile_bytes);
      for (int i = 0; i < p.nwt; ++i) tma_load_3d(w_base + p.wt_off[i], &p.w_map, &w_full, 0, 0, p.wt_widx[i]);
    }
    __syncwarp();
    uint32_t gi = 0;
    for (uint32_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      // 32-bit tile decode (64-bit divisions in every epilogue thread cost a visible share of a tile)
      uint32_t t = tile;
      const uint32_t t1 = t / (uint32_t)p.tiles_w;
      const int tw = (int)(t - t1 * (uint32_t)p.tiles_w);
      const uint32_t t2 = t1 / (uint32_t)p.tiles_h;
      const int th = (int)(t1 - t2 * (uint32_t)p.tiles_h);
      const uint32_t t3 = t2 / (uint32_t)p.tiles_d;
      const int td = (int)(t2 - t3 * (uint32_t)p.tiles_d);
      const int b = (int)t3;
      const int d0 = td * p.bd, h0 = th * p.bh, w0 = tw * p.bw;
      for (int a = 0; a < p.nloads; ++a, ++gi) {
        const uint32_t stage = gi % (uint32_t)p.stages;
        mbar_wait(&empty_bar[stage], ((gi / (uint32_t)p.stages) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], 128u * row_bytes);
          tma_load_5d(a_base + (size_t)stage * p.a_stage_bytes, &p.a_map, &full_bar[stage], 0, w0 + p.load_off[a][2],
                      h0 + p.load_off[a][1], d0 + p.load_off[a][0], b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
    const uint32_t idesc0 = idesc_f16(p.is_f16 != 0, 0u, false, false);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a16 = __shfl_sync(0xffffffffu, (smem_u32(a_base) & 0x3FFFFu) >> 4, 0);
    const uint32_t w16 = __shfl_sync(0xffffffffu, (smem_u32(w_base) & 0x3FFFFu) >> 4, 0);
    const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
    const uint32_t hi = (((8u * row_bytes) >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
    const uint32_t stage16 = (uint32_t)p.a_stage_bytes >> 4;
    const int ksteps = p.KC / 16;
    mbar_wait(&w_full, 0);
    tc_fence_after();
    uint32_t gi = 0, k = 0;
    for (uint32_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++k) {
      const uint32_t buf = k & 1u;
      mbar_wait(&acc_empty[buf], ((k >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t dbase = tmem_u + buf * (uint32_t)p.ncols;
      for (int a = 0; a < p.nloads; ++a, ++gi) {
        const uint32_t stage = gi % (uint32_t)p.stages;
        mbar_wait(&full_bar[stage], (gi / (uint32_t)p.stages) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = a16 + stage * stage16;
          for (int o = p.op_begin[a]; o < p.op_begin[a + 1]; ++o) {
            const uint32_t idesc = idesc0 | (((uint32_t)p.op_n[o] >> 3) << 17);
            const uint32_t sb = w16 + (uint32_t)p.op_w16[o];
            const uint32_t dcol = dbase + (uint32_t)p.op_col[o];
            const uint32_t first = (uint32_t)p.op_first[o];
            for (int ks = 0; ks < ksteps; ++ks)
              umma_f16(dcol, ((uint64_t)hi << 32) | (uint64_t)(sa + 2u * ks), ((uint64_t)hi << 32) | (uint64_t)(sb + 2u * ks),
                       idesc, (first && ks == 0) ? 0u : 1u);
          }
          umma_commit(&empty_bar[stage]);
          if (a == p.nloads - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: set 0 = warps 2..5, set 1 = warps 6..9; thread = one voxel q of the brick (TMEM lane), the groups
    // of the set's d-parity half =====
    const int q = warp & 3;
    const int set = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int qw = row % p.bw, qh = (row / p.bw) % p.bh, qd = row / (p.bw * p.bh);
    const bool issuer = threadIdx.x == 64 + 128 * set;
    uint8_t* o_set = o_base + (size_t)set * p.out_half_bytes;
    if (set < p.os[0]) {
    const int sbw = p.os[2] * p.bw, sbh = p.os[1] * p.bh;  // staged box extents along w, h
    const int nparts = p.hsplit ? p.os[1] : 1;
    const uint32_t orow_bytes = (uint32_t)p.Cout * 2u;
    uint32_t k = 0;
    for (uint32_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++k) {
      // 32-bit tile decode (64-bit divisions in every epilogue thread cost a visible share of a tile)
      uint32_t t = tile;
      const uint32_t t1 = t / (uint32_t)p.tiles_w;
      const int tw = (int)(t - t1 * (uint32_t)p.tiles_w);
      const uint32_t t2 = t1 / (uint32_t)p.tiles_h;
      const int th = (int)(t1 - t2 * (uint32_t)p.tiles_h);
      const uint32_t t3 = t2 / (uint32_t)p.tiles_d;
      const int td = (int)(t2 - t3 * (uint32_t)p.tiles_d);
      const int b = (int)t3;
      const uint32_t buf = k & 1u;
      mbar_wait(&acc_full[buf], (k >> 1) & 1u);
      tc_fence_after();
      const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)p.ncols;
      const int r0 = set;
      for (int r1 = 0; r1 < nparts; ++r1) {
        // the bulk store of the previous part must have finished reading the staging buffer
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
        for (int g = 0; g < p.ngroups; ++g) {
          if (p.grp_r[g][0] != r0 || (p.hsplit && p.grp_r[g][1] != r1)) continue;
          const uint32_t srow = p.hsplit ? (uint32_t)((qd * p.bh + qh) * sbw + (p.os[2] * qw + p.grp_r[g][2]))
                                         : (uint32_t)((qd * sbh + (p.os[1] * qh + p.grp_r[g][1])) * sbw + (p.os[2] * qw + p.grp_r[g][2]));
          if ((p.Cout & 31) == 0) {
            // 32 columns per round: both TMEM loads in flight before the first conversion
            for (int c0 = 0; c0 < p.Cout; c0 += 32) {
              uint32_t ra[16], rb[16];
              tmem_ld16_async(tcol + (uint32_t)(g * p.Cout + c0), ra);
              tmem_ld16_async(tcol + (uint32_t)(g * p.Cout + c0 + 16), rb);
              tmem_ld_fence(ra);
              tmem_ld_fence(rb);
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(h < 2 ? ra[(h & 1) * 8 + j] : rb[(h & 1) * 8 + j]);
                uint32_t off = srow * orow_bytes + (uint32_t)(c0 + 8 * h) * 2u;
                off ^= ((off >> 7) & (uint32_t)p.out_mask) << 4;
                store8<T>(reinterpret_cast<T*>(o_set + off), v);
              }
            }
          } else {
            for (int c0 = 0; c0 < p.Cout; c0 += 16) {
              uint32_t r[16];
              tmem_ld16(tcol + (uint32_t)(g * p.Cout + c0), r);
              float lo[8], hi8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { lo[j] = __uint_as_float(r[j]); hi8[j] = __uint_as_float(r[8 + j]); }
              uint32_t off0 = srow * orow_bytes + (uint32_t)c0 * 2u;
              uint32_t off1 = off0 + 16u;
              off0 ^= ((off0 >> 7) & (uint32_t)p.out_mask) << 4;
              off1 ^= ((off1 >> 7) & (uint32_t)p.out_mask) << 4;
              store8<T>(reinterpret_cast<T*>(o_set + off0), lo);
              store8<T>(reinterpret_cast<T*>(o_set + off1), hi8);
            }
          }
        }
        if (r1 == nparts - 1) {  // every TMEM read of this set's half is done: hand its share of the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
        if (issuer) {
          const int c1 = p.os[2] * tw * p.bw, c2 = (p.hsplit ? 1 : p.os[1]) * th * p.bh, c3 = td * p.bd;
          if (p.accumulate) tma_reduce_add_5d(&p.o_map[2 * r0 + r1], o_set, 0, c1, c2, c3, b);
          else tma_store_5d(&p.o_map[2 * r0 + r1], o_set, 0, c1, c2, c3, b);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

static inline int gm_align1k(long long v) { return (int)(((v + 1023) / 1024) * 1024); }

// Returns MTB200_ERR_UNSUPPORTED when the problem is outside this kernel's envelope (caller falls back).
int conv_gm_umma(const mtb200_conv_params& p, cudaStream_t s) {
  if (p.ngroups < 2 || p.xform || p.bias || p.stats) return MTB200_ERR_UNSUPPORTED;
  if (p.Cin != 16 && p.Cin != 32 && p.Cin != 64) return MTB200_ERR_UNSUPPORTED;           // one K chunk
  if (p.Cout != 16 && p.Cout != 32 && p.Cout != 64) return MTB200_ERR_UNSUPPORTED;        // one swizzled staging row
  int ngexp = 1;
  for (int k = 0; k < 3; ++k) {
    if (p.is[k] != 1 || p.os[k] < 1 || p.os[k] > 2) return MTB200_ERR_UNSUPPORTED;
    ngexp *= p.os[k];
  }
  if (p.ngroups != ngexp || p.ngroups * p.Cout > 256) return MTB200_ERR_UNSUPPORTED;
  if (p.Do * p.os[0] != p.Dof || p.Ho * p.os[1] != p.Hof || p.Wo * p.os[2] != p.Wof) return MTB200_ERR_UNSUPPORTED;
  if (p.ntaps > MTB200_MAX_TAPS) return MTB200_ERR_UNSUPPORTED;

  static thread_local GmParams q;
  memset(&q, 0, sizeof(q));
  // groups must be exactly the residue classes of the output lattice
  bool seen[8] = {false, false, false, false, false, false, false, false};
  for (int g = 0; g < p.ngroups; ++g) {
    int code = 0;
    for (int k = 0; k < 3; ++k) {
      const int r = p.group_ooff[g][k];
      if (r < 0 || r >= p.os[k]) return MTB200_ERR_UNSUPPORTED;
      q.grp_r[g][k] = r;
      code = code * 2 + r;
    }
    if (seen[code]) return MTB200_ERR_UNSUPPORTED;
    seen[code] = true;
  }
  q.KC = p.Cin; q.Cout = p.Cout; q.ngroups = p.ngroups; q.ncols = p.ngroups * p.Cout;
  for (int k = 0; k < 3; ++k) q.os[k] = p.os[k];
  const int rowb = q.KC * 2;
  const int wtile = p.Cout * rowb;

  // distinct input offsets; the one used by the most groups first (it initialises the most accumulator columns)
  int noff = 0;
  int off_of_tap[MTB200_MAX_TAPS], grp_of_tap[MTB200_MAX_TAPS], cnt[MTB200_MAX_GROUPS * 4];
  for (int g = 0; g < p.ngroups; ++g)
    for (int t = p.group_tap_begin[g]; t < p.group_tap_begin[g + 1]; ++t) {
      int a = -1;
      for (int i = 0; i < noff; ++i)
        if (q.load_off[i][0] == p.tap_off[t][0] && q.load_off[i][1] == p.tap_off[t][1] && q.load_off[i][2] == p.tap_off[t][2]) a = i;
      if (a < 0) {
        if (noff >= MTB200_MAX_GROUPS * 4) return MTB200_ERR_UNSUPPORTED;
        a = noff++;
        for (int k = 0; k < 3; ++k) q.load_off[a][k] = p.tap_off[t][k];
        cnt[a] = 0;
      }
      off_of_tap[t] = a; grp_of_tap[t] = g; ++cnt[a];
    }
  // order of the loads: descending group count (stable)
  int order[MTB200_MAX_GROUPS * 4];
  for (int i = 0; i < noff; ++i) order[i] = i;
  for (int i = 1; i < noff; ++i)
    for (int j = i; j > 0 && cnt[order[j]] > cnt[order[j - 1]]; --j) { const int tmp = order[j]; order[j] = order[j - 1]; order[j - 1] = tmp; }
  int sorted_off[MTB200_MAX_GROUPS * 4][3];
  for (int i = 0; i < noff; ++i)
    for (int k = 0; k < 3; ++k) sorted_off[i][k] = q.load_off[order[i]][k];
  // MMA runs: per load, maximal runs of adjacent groups with the same first-touch state
  bool touched[MTB200_MAX_GROUPS];
  for (int g = 0; g < p.ngroups; ++g) touched[g] = false;
  int nops = 0, nwt = 0;
  for (int i = 0; i < noff; ++i) {
    const int a = order[i];
    q.op_begin[i] = nops;
    int widx_of_group[MTB200_MAX_GROUPS];
    for (int g = 0; g < p.ngroups; ++g) widx_of_group[g] = -1;
    for (int t = 0; t < p.ntaps; ++t)
      if (off_of_tap[t] == a) {
        if (widx_of_group[grp_of_tap[t]] >= 0) return MTB200_ERR_UNSUPPORTED;  // two taps of one group at one offset
        widx_of_group[grp_of_tap[t]] = p.tap_widx[t];
      }
    int g = 0;
    while (g < p.ngroups) {
      if (widx_of_group[g] < 0) { ++g; continue; }
      const bool first = !touched[g];
      int e = g;
      while (e < p.ngroups && widx_of_group[e] >= 0 && (!touched[e]) == first) ++e;
      if (nops >= GM_MAX_OPS) return MTB200_ERR_UNSUPPORTED;
      q.op_col[nops] = g * p.Cout;
      q.op_n[nops] = (e - g) * p.Cout;
      q.op_w16[nops] = (nwt * wtile) >> 4;
      q.op_first[nops] = first ? 1 : 0;
      for (int j = g; j < e; ++j) {
        q.wt_widx[nwt] = widx_of_group[j];
        q.wt_off[nwt] = nwt * wtile;
        ++nwt;
        touched[j] = true;
      }
      ++nops;
      g = e;
    }
  }
  q.op_begin[noff] = nops;
  for (int g = 0; g < p.ngroups; ++g)
    if (!touched[g]) return MTB200_ERR_UNSUPPORTED;  // a group without taps would leave its columns uninitialised
  for (int i = 0; i < noff; ++i)
    for (int k = 0; k < 3; ++k) q.load_off[i][k] = sorted_off[i][k];
  q.nloads = noff;
  q.nwt = nwt;

  // brick: powers of two with product 128 minimising the number of tiles (ties: widest in w); staged box dims <= 256
  long long best = -1;
  for (int bw = 128; bw >= 1; bw >>= 1)
    for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
      const int bd = 128 / (bw * bh);
      if (bw * p.os[2] > 256 || bh * p.os[1] > 256 || bd > 256) continue;
      const long long nt = (long long)((p.Do + bd - 1) / bd) * ((p.Ho + bh - 1) / bh) * ((p.Wo + bw - 1) / bw);
      if (best < 0 || nt < best) { best = nt; q.bd = bd; q.bh = bh; q.bw = bw; }
    }
  if (best < 0) return MTB200_ERR_UNSUPPORTED;
  q.tiles_d = (p.Do + q.bd - 1) / q.bd; q.tiles_h = (p.Ho + q.bh - 1) / q.bh; q.tiles_w = (p.Wo + q.bw - 1) / q.bw;
  q.B = p.B;
  if ((long long)p.B * q.tiles_d * q.tiles_h * q.tiles_w >= (1LL << 31) - 65536) return MTB200_ERR_UNSUPPORTED;
  q.ntiles = (uint32_t)((long long)p.B * q.tiles_d * q.tiles_h * q.tiles_w);
  q.a_stage_bytes = 128 * rowb;
  q.w_bytes = gm_align1k((long long)nwt * wtile);
  const int orowb = p.Cout * 2;
  // one staging buffer per d-parity half (epilogue set): the whole half brick when that leaves >= 4 load stages (one
  // store per set and tile: transposed conv 0.47 -> 0.27 ms), else one h-parity quarter at a time (strided dgrad: 27
  // resident weight tiles)
  q.out_mask = orowb == 128 ? 7 : (orowb == 64 ? 3 : 1);
  const int budget = 226 * 1024;
  int fixed = 0;
  static const int gm_dbg = [] { const char* e = getenv("MTB200_GM_DBG"); return e ? atoi(e) : 0; }();  // bit 1: quarters always
  for (q.hsplit = (gm_dbg & 2) ? 1 : 0; q.hsplit < 2; ++q.hsplit) {
    q.out_half_bytes = gm_align1k((long long)q.bd * ((q.hsplit ? 1 : p.os[1]) * q.bh) * (p.os[2] * q.bw) * orowb);
    fixed = q.w_bytes + p.os[0] * q.out_half_bytes + 1024;
    q.stages = min(GM_MAX_STAGES, (budget - fixed) / q.a_stage_bytes);
    if (q.stages >= 4 || p.os[1] == 1) break;
  }
  if (q.hsplit > 1) q.hsplit = 1;
  if (q.stages < 2) return MTB200_ERR_UNSUPPORTED;
  q.accumulate = p.accumulate;
  {  // experiments only (timing, wrong results): MTB200_GM_DBG=1 stores instead of reduce-adding
    static const int dbg = [] { const char* e = getenv("MTB200_GM_DBG"); return e ? atoi(e) : 0; }();
    if (dbg & 1) q.accumulate = 0;
  }
  q.is_f16 = p.dtype == MTB200_F16;
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.Di, (cuuint64_t)p.B};
    cuuint64_t strides[4] = {(cuuint64_t)p.in_ldc * 2, (cuuint64_t)p.Wi * p.in_ldc * 2,
                             (cuuint64_t)p.Hi * p.Wi * p.in_ldc * 2, (cuuint64_t)p.Di * p.Hi * p.Wi * p.in_ldc * 2};
    cuuint32_t box[5] = {(cuuint32_t)q.KC, (cuuint32_t)q.bw, (cuuint32_t)q.bh, (cuuint32_t)q.bd, 1};
    if (!umma_encode_map(&q.a_map, p.dtype, 5, (uint8_t*)p.in + (size_t)p.in_coff * 2, dims, strides, box, rowb))
      return MTB200_ERR_CUDA;
  }
  {
    int n_widx = 0;
    for (int t = 0; t < p.ntaps; ++t) n_widx = max(n_widx, p.tap_widx[t] + 1);
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, (cuuint64_t)n_widx};
    cuuint64_t strides[2] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)p.Cin * p.Cout * 2};
    cuuint32_t box[3] = {(cuuint32_t)q.KC, (cuuint32_t)p.Cout, 1};
    if (!umma_encode_map(&q.w_map, p.dtype, 3, (void*)p.w, dims, strides, box, rowb)) return MTB200_ERR_CUDA;
  }
  for (int r0 = 0; r0 < p.os[0]; ++r0)
    for (int r1 = 0; r1 < (q.hsplit ? p.os[1] : 1); ++r1) {
      const int hs = q.hsplit ? p.os[1] : 1;  // h stride of the stored sub-lattice
      // output voxels with d = os0 * qd + r0, h = os1 * qh + r1: base shifted by r0 planes and r1 lines, the d and h
      // strides multiplied by os0, os1
      const long long ext_d = (p.Dof - r0 + p.os[0] - 1) / p.os[0];
      const long long ext_h = (p.Hof - r1 + hs - 1) / hs;
      cuuint64_t dims[5] = {(cuuint64_t)p.Cout, (cuuint64_t)p.Wof, (cuuint64_t)ext_h, (cuuint64_t)ext_d, (cuuint64_t)p.B};
      cuuint64_t strides[4] = {(cuuint64_t)p.out_ldc * 2, (cuuint64_t)p.Wof * p.out_ldc * 2 * hs,
                               (cuuint64_t)p.Hof * p.Wof * p.out_ldc * 2 * p.os[0],
                               (cuuint64_t)p.Dof * p.Hof * p.Wof * p.out_ldc * 2};
      cuuint32_t box[5] = {(cuuint32_t)p.Cout, (cuuint32_t)(p.os[2] * q.bw), (cuuint32_t)((p.os[1] / hs) * q.bh), (cuuint32_t)q.bd, 1};
      uint8_t* base = (uint8_t*)p.out + (((size_t)r0 * p.Hof + r1) * p.Wof * p.out_ldc + p.out_coff) * 2;
      if (!umma_encode_map(&q.o_map[2 * r0 + r1], p.dtype, 5, base, dims, strides, box, orowb)) return MTB200_ERR_CUDA;
    }
  const int smem = q.stages * q.a_stage_bytes + fixed;
  const int gx = (int)min((long long)q.ntiles, (long long)num_sms());
  cudaError_t e;
  if (p.dtype == MTB200_BF16) {
    e = cudaFuncSetAttribute(conv_gm_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) launch_pdl(conv_gm_umma_kernel<__nv_bfloat16>, dim3(gx), dim3(GM_THREADS), (size_t)(smem), s, q);
  } else {
    e = cudaFuncSetAttribute(conv_gm_umma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) launch_pdl(conv_gm_umma_kernel<__half>, dim3(gx), dim3(GM_THREADS), (size_t)(smem), s, q);
  }
  if (e != cudaSuccess) { set_error("conv_gm: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  return check_launch("conv_gm_umma");
}

}  // namespace mtb
