// Plane-streaming ("halo-resident", input-plane-stationary) tcgen05 convolution for stride-1 kernels with taps in
// [-1,1]^3 -- Conv3d 3x3x3 / 1x3x3 forward and its data gradient (generic_UNet.py:46, the 3x3x3 stack).
//
// The per-tap kernel (conv_umma.cu) re-fetches every activation tile once per tap: 27x L2->SM traffic, which saturates
// the L2 path long before the tensor pipe.  Here a CTA owns an output column
//     16 h-lines x (8*TW) w-columns x a segment [ds, de) of d-planes
// and streams the INPUT planes through a small ring in shared memory exactly once: plane p = one 5-D TMA box per 64-
// channel chunk, [18 h][8*TW+2 w][chunk] with the halo included (out-of-volume = zero fill = the conv's padding).  A tap
// (dz,dy,dx) is a shifted VIEW of that plane: the K-major UMMA descriptor starts at row (dy+1)*Wh + (dx+1) + 8*tw and
// steps Wh rows between 8-row groups (tools/umma_probe.cu: any row-granular start / group stride works with
// base_offset = 0 because the swizzle is a function of the shared-memory address).
//
// Input-plane-stationary order: when plane p has landed, ALL its taps are issued -- dz=+1 into the accumulator of output
// plane p-1 (which completes it), dz=0 into plane p, dz=-1 into plane p+1 (which opens it) -- so a plane is dead after
// one visit (ring of 2-3 planes instead of 4) and three accumulator sets of TW x BN TMEM columns rotate; the epilogue of
// plane p-1 overlaps the remaining two thirds of plane p's MMAs.  Weight tiles [BN][Cin] per tap stay resident in
// shared memory when all of them fit, else they stream through a ring in consumption order.
//
// Warp roles (7 warps): 0 = plane producer (TMA), 1 = TMEM owner + MMA issuer (one thread, descriptor low words are
// plain 32-bit adds), 2 = weight producer (TMA), 3..6 = epilogue (TMEM -> +bias -> round -> global, InstanceNorm
// sum / sum-of-squares kept in registers per thread and reduced across the warp once per CTA).
#include "umma.cuh"

namespace mtb {

using namespace um;

constexpr int HL_THREADS = 224;
constexpr int HL_MAX_RING = 4;
constexpr int HL_MAX_WSTAGES = 16;
constexpr int HL_HT = 16;  // output h-lines per CTA
constexpr int HL_HH = HL_HT + 2;
constexpr int HL_NSETS = 3;

struct HaloParams {
  CUtensorMap a_map, w_map;
  void* out;
  const float* bias;
  double* stats;
  int B, D, H, W;  // output grid == input grid (stride 1)
  int out_ldc, out_coff, Cout;
  int TW, Wh, BN;
  int kcw, nchunk;              // channels per chunk (<= 64), chunks per voxel row
  int sub_bytes, plane_bytes;   // one chunk's sub-plane (1024-aligned), one plane (= nchunk sub-planes)
  int plane_tx, ring;
  int wchunk_bytes, wtile_bytes;  // one (tap, chunk) weight tile [BN][kcw] (1024-aligned), one tap (= nchunk tiles)
  int w_resident, wstages;
  int tmem_cols;
  int tiles_h, tiles_w, nseg, seglen;
  int dzmin, dzmax;
  int grp_begin[4];             // taps sorted by dz = +1, 0, -1
  int tap_rowoff[MTB200_MAX_TAPS];  // ((dy+1)*Wh + dx+1) rows
  int tap_widx[MTB200_MAX_TAPS];
  int ntaps;
  int accumulate, is_f16;
};

__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | (uint64_t)lo; }

// high word of a K-major descriptor: SBO (bits 32..45), version (bit 46), swizzle mode (bits 61..63)
__device__ __forceinline__ uint32_t kmajor_hi(uint32_t row_bytes, uint32_t sbo) {
  const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
  return ((sbo >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
}

template <typename T, int ROWB>
__global__ void __launch_bounds__(HL_THREADS, 1) conv_halo_umma_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t plane_full[HL_MAX_RING], plane_empty[HL_MAX_RING];
  __shared__ __align__(8) uint64_t w_full[HL_MAX_WSTAGES], w_empty[HL_MAX_WSTAGES];
  __shared__ __align__(8) uint64_t acc_full[HL_NSETS], acc_empty[HL_NSETS];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[64], s_sum[4][64], s_sq[4][64];  // statistics: one slot per epilogue warp, no float atomics

  constexpr int KSTEPS = ROWB / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* plane_base = dsmem;
  uint8_t* w_base = dsmem + (size_t)p.ring * p.plane_bytes;

  // work unit
  int u = blockIdx.x;
  const int seg = u % p.nseg; u /= p.nseg;
  const int twg = u % p.tiles_w; u /= p.tiles_w;
  const int th = u % p.tiles_h;
  const int b = u / p.tiles_h;
  const int h0 = th * HL_HT, w0 = twg * 8 * p.TW;
  const int ds = seg * p.seglen, de = min(p.D, ds + p.seglen);  // >= 1 output plane by construction
  const int n0 = blockIdx.y * p.BN;
  const int pfirst = max(0, ds + p.dzmin), plast = min(p.D - 1, de - 1 + p.dzmax);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.ring; ++i) { mbar_init(&plane_full[i], 1); mbar_init(&plane_empty[i], 1); }
    for (int i = 0; i < p.wstages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < HL_NSETS; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 64) {
    for (int w = 0; w < 4; ++w) { s_sum[w][threadIdx.x] = 0.f; s_sq[w][threadIdx.x] = 0.f; }
    s_bias[threadIdx.x] = (p.bias && (int)threadIdx.x < p.BN) ? p.bias[n0 + threadIdx.x] : 0.f;
  }
  if (warp == 1) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== input-plane producer (warp-uniform loop, one elected lane issues the TMA) =====
    {
      int j = 0;
      for (int pl = pfirst; pl <= plast; ++pl, ++j) {
        const int slot = j % p.ring;
        mbar_wait(&plane_empty[slot], (((uint32_t)(j / p.ring)) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&plane_full[slot], (uint32_t)p.plane_tx);
          for (int c = 0; c < p.nchunk; ++c)
            tma_load_5d(plane_base + (size_t)slot * p.plane_bytes + (size_t)c * p.sub_bytes, &p.a_map,
                        &plane_full[slot], c * p.kcw, w0 - 1, h0 - 1, pl, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ===== weight producer =====
    if (p.w_resident) {
      if (elect_one()) {
        mbar_expect_tx(&w_full[0], (uint32_t)(p.ntaps * p.nchunk * p.BN * ROWB));
        for (int t = 0; t < p.ntaps; ++t)
          for (int c = 0; c < p.nchunk; ++c)
            tma_load_3d(w_base + (size_t)t * p.wtile_bytes + (size_t)c * p.wchunk_bytes, &p.w_map, &w_full[0],
                        c * p.kcw, n0, p.tap_widx[t]);
      }
      __syncwarp();
    } else {
      uint32_t wc = 0;
      for (int pl = pfirst; pl <= plast; ++pl)
        for (int gi = 0; gi < 3; ++gi) {
          const int o = pl - (1 - gi);
          if (o < ds || o >= de) continue;
          for (int t = p.grp_begin[gi]; t < p.grp_begin[gi + 1]; ++t, ++wc) {
            const uint32_t ws = wc % (uint32_t)p.wstages;
            mbar_wait(&w_empty[ws], ((wc / (uint32_t)p.wstages) & 1u) ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(&w_full[ws], (uint32_t)(p.nchunk * p.BN * ROWB));
              for (int c = 0; c < p.nchunk; ++c)
                tma_load_3d(w_base + (size_t)ws * p.wtile_bytes + (size_t)c * p.wchunk_bytes, &p.w_map, &w_full[ws],
                            c * p.kcw, n0, p.tap_widx[t]);
            }
            __syncwarp();
          }
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the (warp-uniform) loop, one elected lane issues =====
    {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc = idesc_f16(p.is_f16 != 0, (uint32_t)p.BN, false, false);
      const uint32_t hi_a = kmajor_hi(ROWB, (uint32_t)(p.Wh * ROWB)), hi_b = kmajor_hi(ROWB, 8u * ROWB);
      const uint32_t sub16 = (uint32_t)p.sub_bytes >> 4, wchunk16 = (uint32_t)p.wchunk_bytes >> 4;
      const uint32_t w16 = __shfl_sync(0xffffffffu, (smem_u32(w_base) & 0x3FFFFu) >> 4, 0);
      const uint32_t wtile16 = (uint32_t)p.wtile_bytes >> 4;
      const uint32_t plane16 = __shfl_sync(0xffffffffu, (smem_u32(plane_base) & 0x3FFFFu) >> 4, 0);
      const uint32_t pstride16 = (uint32_t)p.plane_bytes >> 4;
      const int TW = p.TW, nchunk = p.nchunk;
      uint32_t wc = 0;
      if (p.w_resident) { mbar_wait(&w_full[0], 0); tc_fence_after(); }
      int j = 0;
      for (int pl = pfirst; pl <= plast; ++pl, ++j) {
        const int slot = j % p.ring;
        mbar_wait(&plane_full[slot], ((uint32_t)(j / p.ring)) & 1u);
        tc_fence_after();
        const uint32_t a_plane = plane16 + (uint32_t)slot * pstride16;
        for (int gi = 0; gi < 3; ++gi) {
          const int dz = 1 - gi;
          const int o = pl - dz;
          const int tb = p.grp_begin[gi], te = p.grp_begin[gi + 1];
          if (o < ds || o >= de || tb == te) continue;
          const int rel = o - ds;
          const int set = rel % HL_NSETS;
          uint32_t accf = 1u;
          if (pl == max(o + p.dzmin, 0)) {  // first contribution: the set must have been drained by the epilogue
            mbar_wait(&acc_empty[set], (((uint32_t)(rel / HL_NSETS)) & 1u) ^ 1u);
            tc_fence_after();
            accf = 0u;
          }
          const uint32_t dset = tmem_u + (uint32_t)(set * TW * p.BN);
          for (int t = tb; t < te; ++t) {
            uint32_t b_t;
            uint32_t ws = 0;
            if (p.w_resident) {
              b_t = w16 + (uint32_t)t * wtile16;
            } else {
              ws = wc % (uint32_t)p.wstages;
              mbar_wait(&w_full[ws], (wc / (uint32_t)p.wstages) & 1u);
              tc_fence_after();
              b_t = w16 + ws * wtile16;
              ++wc;
            }
            const uint32_t a_t = a_plane + (uint32_t)(p.tap_rowoff[t] * (ROWB / 16));
            const uint32_t acc_t = (t > tb) ? 1u : accf;
            if (elect_one()) {
              for (int tw = 0; tw < TW; ++tw) {
                const uint32_t dcol = dset + (uint32_t)(tw * p.BN);
                const uint32_t a_tw = a_t + (uint32_t)(tw * 8 * (ROWB / 16));
                for (int c = 0; c < nchunk; ++c) {
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k)
                    umma_f16(dcol, desc64(hi_a, a_tw + (uint32_t)c * sub16 + (uint32_t)(k * 2)),
                             desc64(hi_b, b_t + (uint32_t)c * wchunk16 + (uint32_t)(k * 2)), idesc,
                             (c | k) ? 1u : acc_t);
                }
              }
              if (!p.w_resident) umma_commit(&w_empty[ws]);
            }
            __syncwarp();
          }
          if (pl == min(o + p.dzmax, p.D - 1)) {  // output plane o is complete
            if (elect_one()) umma_commit(&acc_full[set]);
            __syncwarp();
          }
        }
        if (elect_one()) umma_commit(&plane_empty[slot]);
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue warps 3..6 (TMEM lane quarter = warp % 4) =====
    const int q = warp & 3;
    const int m = q * 32 + lane;  // row of the 128-row tile
    const int hh = h0 + (m >> 3);
    T* out = reinterpret_cast<T*>(p.out);
    float csum[4][16], csq[4][16];  // per-thread column partials over every plane of this CTA
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int j = 0; j < 16; ++j) { csum[c][j] = 0.f; csq[c][j] = 0.f; }
    const bool want_stats = p.stats != nullptr;
    for (int o = ds; o < de; ++o) {
      const int rel = o - ds;
      const int set = rel % HL_NSETS;
      mbar_wait(&acc_full[set], ((uint32_t)(rel / HL_NSETS)) & 1u);
      tc_fence_after();
      for (int tw = 0; tw < p.TW; ++tw) {
        const int ww = w0 + tw * 8 + (m & 7);
        const bool valid = hh < p.H && ww < p.W;
        T* orow = out + ((((long long)b * p.D + o) * p.H + hh) * p.W + ww) * p.out_ldc + p.out_coff + n0;
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((set * p.TW + tw) * p.BN);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int c0 = cc * 16;
          if (c0 < p.BN) {
            uint32_t r[16];
            float v[16];
            tmem_ld16(tcol + (uint32_t)c0, r);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + s_bias[c0 + j];
            if (valid) {
              if (p.accumulate) {
                float ov[8];
                load8<T>(orow + c0, ov);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += ov[j];
                load8<T>(orow + c0 + 8, ov);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[8 + j] += ov[j];
              }
              float lo[8], hi[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
              store8<T>(orow + c0, lo);
              store8<T>(orow + c0 + 8, hi);
              if (want_stats) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float x = Traits<T>::round(v[j]);
                  csum[cc][j] += x;
                  csq[cc][j] = fmaf(x, x, csq[cc][j]);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[set]);
    }
    if (want_stats) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
        if (cc * 16 < p.BN) {
          warp_colsum16(csum[cc], lane);
          warp_colsum16(csq[cc], lane);
          if ((lane & 1) == 0) {
            const int col = colsum16_column(lane);
            s_sum[q][cc * 16 + col] = csum[cc][0];  // once per CTA, one owner lane per (warp, column)
            s_sq[q][cc * 16 + col] = csq[cc][0];
          }
        }
    }
  }
  __syncthreads();
  if (p.stats) {
    for (int c = threadIdx.x; c < p.BN; c += HL_THREADS) {
      const float su = ((s_sum[0][c] + s_sum[1][c]) + s_sum[2][c]) + s_sum[3][c];
      const float sq = ((s_sq[0][c] + s_sq[1][c]) + s_sq[2][c]) + s_sq[3][c];
      if (su != 0.f || sq != 0.f) {
        double* st = p.stats + ((long long)b * p.Cout + n0 + c) * 2;
        atomicAdd(st, (double)su);
        atomicAdd(st + 1, (double)sq);
      }
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <typename T, int ROWB>
static cudaError_t launch_halo(const HaloParams& q, dim3 grid, int smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(conv_halo_umma_kernel<T, ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) conv_halo_umma_kernel<T, ROWB><<<grid, HL_THREADS, smem, s>>>(q);
  return e;
}

static inline int align1k(long long v) { return (int)(((v + 1023) / 1024) * 1024); }

// Returns MTB200_ERR_UNSUPPORTED when the problem is outside this kernel's envelope (caller falls back to the per-tap
// kernel); all generic argument checks were done by the caller.
int conv_halo_umma(const mtb200_conv_params& p, cudaStream_t s) {
  if (p.ngroups != 1) return MTB200_ERR_UNSUPPORTED;
  for (int k = 0; k < 3; ++k)
    if (p.is[k] != 1 || p.os[k] != 1 || p.group_ooff[0][k] != 0) return MTB200_ERR_UNSUPPORTED;
  if (p.Do != p.Di || p.Ho != p.Hi || p.Wo != p.Wi || p.Dof != p.Do || p.Hof != p.Ho || p.Wof != p.Wo)
    return MTB200_ERR_UNSUPPORTED;
  if (p.Cin != 16 && p.Cin != 32 && p.Cin % 64 != 0) return MTB200_ERR_UNSUPPORTED;
  bool spatial = false, have_dz0 = false;
  for (int t = 0; t < p.ntaps; ++t) {
    for (int k = 0; k < 3; ++k) {
      if (p.tap_off[t][k] < -1 || p.tap_off[t][k] > 1) return MTB200_ERR_UNSUPPORTED;
      if (p.tap_off[t][k] != 0) spatial = true;
    }
    if (p.tap_off[t][0] == 0) have_dz0 = true;
  }
  if (!spatial || !have_dz0 || p.ntaps < 9) return MTB200_ERR_UNSUPPORTED;  // 1x1x1: nothing to reuse
  if (p.Ho < 8 || p.Wo < 8) return MTB200_ERR_UNSUPPORTED;                  // tiny maps: tiles would be mostly padding

  static thread_local HaloParams q;
  memset(&q, 0, sizeof(q));
  q.kcw = p.Cin < 64 ? p.Cin : 64;
  q.nchunk = p.Cin / q.kcw;
  const int rowb = q.kcw * 2;

  // ---- configuration search: N tile, w sub-tiles, resident / streamed weights, plane ring.
  // Cost model = L2->SM bytes per tensor-pipe cycle (weights re-streamed per plane unless resident + the halo planes,
  // fetched by each of the Cout/BN CTAs of a column); the per-SM L2 path sustains ~40 B/cycle.
  const int smem_budget = 222 * 1024;
  double best_cost = 1e30;
  int best_bn = 0, best_tw = 0, best_res = 0, best_ring = 0, best_ws = 0;
  for (int bn = 64; bn >= 16; bn >>= 1) {
    if (p.Cout % bn) continue;
    const int wchunk = align1k((long long)bn * rowb);
    const int wtile = wchunk * q.nchunk;
    for (int tw = 4; tw >= 1; --tw) {
      if (HL_NSETS * tw * bn > 512) continue;
      if (tw > 1 && 8 * (tw - 1) >= p.Wo) continue;  // do not tile wider than the map needs
      const int wh = 8 * tw + 2;
      const int sub = align1k((long long)HL_HH * wh * rowb);
      const int plane = sub * q.nchunk;
      for (int res = 1; res >= 0; --res) {
        const long long wbytes = res ? (long long)wtile * p.ntaps : 0;
        for (int ring = 3; ring >= 2; --ring) {
          long long left = smem_budget - wbytes - (long long)ring * plane;
          int ws = 1;
          if (!res) {
            ws = (int)min((long long)HL_MAX_WSTAGES, left / wtile);
            if (ws < 3) continue;
          } else if (left < 0) {
            continue;
          }
          // per plane of 128*tw voxels.  tcgen05.mma (M = 128, K = 16) costs max(N/2, 34.6 + 0.2375 N) cycles
          // (profiles/r1h_umma_rate_probe.txt): narrow N tiles are operand-fetch bound, so N = 32 is NOT half of N = 64
          const double mma_one = fmax(bn / 2.0, 34.6 + 0.2375 * bn);
          const double mma_cycles = (double)p.ntaps * tw * (p.Cin / 16) * mma_one;
          const double wtraffic = res ? 0.0 : (double)p.ntaps * bn * p.Cin * 2;
          const double atraffic = (double)HL_HH * wh * p.Cin * 2;
          // cycles per output voxel and Cout: the slower of the tensor pipe and the L2->SM path (~40 B/cycle/SM); the
          // Cout/bn CTAs of a column each stream the planes again
          double cost = fmax(mma_cycles, (wtraffic + atraffic) / 40.0) / (128.0 * tw) * (p.Cout / bn);
          if (ring == 2) cost *= 1.05;  // less prefetch distance
          // tile-quantisation waste in w
          const int tiles_w = (p.Wo + 8 * tw - 1) / (8 * tw);
          cost *= (double)(tiles_w * 8 * tw) / p.Wo;
          if (cost < best_cost - 1e-9) {
            best_cost = cost; best_bn = bn; best_tw = tw; best_res = res; best_ring = ring; best_ws = ws;
          }
        }
      }
    }
  }
  if (best_bn == 0) return MTB200_ERR_UNSUPPORTED;
  q.BN = best_bn; q.TW = best_tw; q.w_resident = best_res; q.ring = best_ring; q.wstages = best_ws;
  q.Wh = 8 * q.TW + 2;
  q.sub_bytes = align1k((long long)HL_HH * q.Wh * rowb);
  q.plane_bytes = q.sub_bytes * q.nchunk;
  q.plane_tx = q.nchunk * HL_HH * q.Wh * rowb;
  q.wchunk_bytes = align1k((long long)q.BN * rowb);
  q.wtile_bytes = q.wchunk_bytes * q.nchunk;
  q.tmem_cols = 32;
  while (q.tmem_cols < HL_NSETS * q.TW * q.BN) q.tmem_cols *= 2;

  // taps sorted by dz = +1, 0, -1
  {
    int n = 0;
    q.dzmin = 0; q.dzmax = 0;
    for (int gi = 0; gi < 3; ++gi) {
      const int dz = 1 - gi;
      q.grp_begin[gi] = n;
      for (int t = 0; t < p.ntaps; ++t)
        if (p.tap_off[t][0] == dz) {
          q.tap_rowoff[n] = (p.tap_off[t][1] + 1) * q.Wh + p.tap_off[t][2] + 1;
          q.tap_widx[n] = p.tap_widx[t];
          ++n;
          q.dzmin = min(q.dzmin, dz); q.dzmax = max(q.dzmax, dz);
        }
    }
    q.grp_begin[3] = n;
    q.ntaps = n;
  }

  // tensor maps: activations [C][W][H][D][B] (box = one chunk of one halo plane), weights [Cin][Cout][taps]
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.Di, (cuuint64_t)p.B};
    cuuint64_t strides[4] = {(cuuint64_t)p.in_ldc * 2, (cuuint64_t)p.Wi * p.in_ldc * 2,
                             (cuuint64_t)p.Hi * p.Wi * p.in_ldc * 2, (cuuint64_t)p.Di * p.Hi * p.Wi * p.in_ldc * 2};
    cuuint32_t box[5] = {(cuuint32_t)q.kcw, (cuuint32_t)q.Wh, (cuuint32_t)HL_HH, 1, 1};
    if (!umma_encode_map(&q.a_map, p.dtype, 5, (uint8_t*)p.in + (size_t)p.in_coff * 2, dims, strides, box, rowb))
      return MTB200_ERR_CUDA;
  }
  {
    int n_widx = 0;
    for (int t = 0; t < p.ntaps; ++t) n_widx = max(n_widx, p.tap_widx[t] + 1);
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, (cuuint64_t)n_widx};
    cuuint64_t strides[2] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)p.Cin * p.Cout * 2};
    cuuint32_t box[3] = {(cuuint32_t)q.kcw, (cuuint32_t)q.BN, 1};
    if (!umma_encode_map(&q.w_map, p.dtype, 3, (void*)p.w, dims, strides, box, rowb)) return MTB200_ERR_CUDA;
  }
  q.out = p.out; q.bias = p.bias; q.stats = p.stats;
  q.B = p.B; q.D = p.Do; q.H = p.Ho; q.W = p.Wo;
  q.out_ldc = p.out_ldc; q.out_coff = p.out_coff; q.Cout = p.Cout;
  q.accumulate = p.accumulate;
  q.is_f16 = p.dtype == MTB200_F16;
  q.tiles_h = (p.Ho + HL_HT - 1) / HL_HT;
  q.tiles_w = (p.Wo + 8 * q.TW - 1) / (8 * q.TW);
  const int ny = p.Cout / q.BN;
  const long long cols = (long long)p.B * q.tiles_h * q.tiles_w;
  // split D into segments: fill whole waves of the machine (one CTA per SM); every segment re-loads its halo planes
  {
    const int sms = num_sms();
    double best = -1;
    int best_nseg = 1;
    const int max_nseg = max(1, p.Do / 4);
    for (int nseg = 1; nseg <= max_nseg; ++nseg) {
      const int seglen = (p.Do + nseg - 1) / nseg;
      const int nreal = (p.Do + seglen - 1) / seglen;
      if (nreal != nseg) continue;
      const long long ctas = cols * nseg * ny;
      const long long waves = (ctas + sms - 1) / sms;
      const double eff = (double)ctas / (double)(waves * sms) * seglen / (seglen + 2.5);  // halo planes + pipeline fill
      if (eff > best) { best = eff; best_nseg = nseg; }
    }
    q.nseg = best_nseg;
    q.seglen = (p.Do + q.nseg - 1) / q.nseg;
  }
  const long long units = cols * q.nseg;
  MTB_REQUIRE(units < (1LL << 31), "conv_halo: too many work units");

  // >= 116 KB so that two CTAs never share an SM (each allocates up to all 512 TMEM columns)
  const int smem = max(116 * 1024, q.ring * q.plane_bytes + (q.w_resident ? q.ntaps : q.wstages) * q.wtile_bytes + 1024);
  dim3 grid((unsigned)units, ny, 1);
  cudaError_t e;
  if (p.dtype == MTB200_BF16) {
    e = rowb == 128 ? launch_halo<__nv_bfloat16, 128>(q, grid, smem, s)
                    : (rowb == 64 ? launch_halo<__nv_bfloat16, 64>(q, grid, smem, s)
                                  : launch_halo<__nv_bfloat16, 32>(q, grid, smem, s));
  } else {
    e = rowb == 128 ? launch_halo<__half, 128>(q, grid, smem, s)
                    : (rowb == 64 ? launch_halo<__half, 64>(q, grid, smem, s) : launch_halo<__half, 32>(q, grid, smem, s));
  }
  if (e != cudaSuccess) { set_error("conv_halo: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  return check_launch("conv_halo_umma");
}

}  // namespace mtb
