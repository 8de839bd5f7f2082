// Line-streaming tcgen05 weight gradient for the WIDE 3x3x3 layers below the second level (stride 1, Cin and Cout
// multiples of 128, 16..64-voxel lines: 128->128 / 256->128 at 48x40x32, 256->256 / 512->256 at 24x20x16):
//     dW[(dz,dy,dx)][co][ci] = sum_{b,d,h,w} dY[b,d,h,w][co] * X[b,d+dz,h+dy,w+dx][ci]
//
// A 128 x 128 block of ONE tap fills a quarter of TMEM, so a CTA can hold four taps at most and every byte it loads can
// feed at most 4 x 128 MACs per element: the per-tap kernel loaded FOUR tap-shifted X bricks + one dY brick per 16 MMAs
// (80 KB per 1024 clk, twice what TMA delivers to an SM: 360..460 TFLOP/s), and wgrad_line's 32 x 32 blocks re-stream both
// tensors once per (Cin chunk, Cout block) pair (16..32 pairs here).  This kernel gives a CTA ONE (dz, dy) pair, one
// 128-channel Cin block and one 128-channel Cout block:
//   A = the X line (b, d+dz, h+dy, -1 .. W) [K = w][M = 128 ci], MN-major straight from NDHWC (two 64-channel SWIZZLE_128B
//       boxes); the three dx taps are ROW-SHIFTED views of that one line (start row dx + 1) -> 3 accumulators in TMEM
//   B = the dY line (b, d, h, 0 .. W) [K = w][N = 128 co], MN-major (two boxes)
//   per line: W/16 K steps x 3 taps MMAs (N = 128: the tensor pipe's full rate) for (W + 2 + W) x 256 B of TMA traffic
//   = 43 B/clk at W = 32 -- the L2 -> SM limit and the MMA rate meet, where the per-tap kernel needed 80.
// grid = 9 (dz, dy) groups x (Cin / 128) x (Cout / 128) x K-split slots (contiguous line ranges); lines whose X line lies
// outside the volume contribute nothing and are skipped.  Epilogue once per CTA: fp32 atomics into dW.
//
// Warp roles (6 warps): 0 = producer (TMA), 1 = TMEM owner + MMA issuer, 2..5 = epilogue.
#include "umma.cuh"

namespace mtb {

using namespace um;

constexpr int WR_THREADS = 192;
constexpr int WR_MAX_STAGES = 8;

struct WgradRowsParams {
  CUtensorMap x_map, dy_map;
  float* dw;
  int B, D, H, W;
  int Cin, Cout;
  int nci, nco;          // 128-channel blocks
  int nkk;               // K steps per line (W / 16)
  int lps;               // h lines per pipeline step (2 on 16-voxel lines: one line alone is three MMAs per barrier round)
  int xrows;             // rows of one staged X line (W + 2)
  int xbox_bytes, ybox_bytes, stage_bytes, stage_tx, stages;
  int nslots;            // K-split slots per (group, ci block, co block)
  int lines, lines_per_slot;
  int lut[27];           // [dz+1][dy+1][dx+1] -> weight slice or -1
  int is_f16;
};

__device__ __forceinline__ uint64_t wr_desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | (uint64_t)lo; }

__global__ void __launch_bounds__(WR_THREADS, 1) wgrad_rows_umma_kernel(const __grid_constant__ WgradRowsParams p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t st_full[WR_MAX_STAGES], st_empty[WR_MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_any;  // the MMA warp issued at least one line

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  // CTA -> (slot, co block, ci block, group); the 9 groups of one (slot, pair) are neighbours: they read the same lines
  int id = (int)blockIdx.x;
  const int g = id % 9; id /= 9;
  const int cib = id % p.nci; id /= p.nci;
  const int cob = id % p.nco; id /= p.nco;
  const int slot = id;
  const int dz = g / 3 - 1, dy = g % 3 - 1;
  const int l0 = slot * p.lines_per_slot;
  const int l1 = min(p.lines, l0 + p.lines_per_slot);

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

  // line l = (b, d, h), walked with counters (no division per line); a line takes part iff its X line (d + dz, h + dy)
  // is inside the volume -- the same predicate in the producer and in the MMA issuer
  // (with lps > 1 a "line" is a group of lps consecutive h lines; p.H counts groups)
  int h0 = l0 % p.H, d0 = (l0 / p.H) % p.D, b0 = l0 / (p.H * p.D);
  const int lps = p.lps;

  if (warp == 0) {
    // ===== producer =====
    uint32_t sc = 0, st = 0, ph = 1;
    int b = b0, d = d0, h = h0;
    for (int l = l0; l < l1; ++l) {
      const bool act = (unsigned)(d + dz) < (unsigned)p.D && (lps > 1 || (unsigned)(h + dy) < (unsigned)p.H);
      const int hh = h * lps, dd = d, bb = b;
      if (++h == p.H) { h = 0; if (++d == p.D) { d = 0; ++b; } }
      if (!act) continue;
      mbar_wait(&st_empty[st], ph);
      if (elect_one()) {
        uint8_t* dst = dsmem + (size_t)st * p.stage_bytes;
        mbar_expect_tx(&st_full[st], (uint32_t)p.stage_tx);
        tma_load_5d(dst, &p.x_map, &st_full[st], cib * 128, -1, hh + dy, dd + dz, bb);
        tma_load_5d(dst + p.xbox_bytes, &p.x_map, &st_full[st], cib * 128 + 64, -1, hh + dy, dd + dz, bb);
        tma_load_5d(dst + 2 * p.xbox_bytes, &p.dy_map, &st_full[st], cob * 128, 0, hh, dd, bb);
        tma_load_5d(dst + 2 * p.xbox_bytes + p.ybox_bytes, &p.dy_map, &st_full[st], cob * 128 + 64, 0, hh, dd, bb);
      }
      __syncwarp();
      ++sc;
      if (++st == (uint32_t)p.stages) { st = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t fmt = p.is_f16 ? 0u : 1u;
    // D = f32, A/B 16-bit, both MN-major, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) |
                           ((128u >> 4) << 24);
    // MN-major SWIZZLE_128B: SBO = 8 rows of 128 B; LBO = distance between the two 64-channel boxes
    const uint32_t hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t lbo_a = ((uint32_t)p.xbox_bytes >> 4) << 16, lbo_b = ((uint32_t)p.ybox_bytes >> 4) << 16;
    const uint32_t s16 = __shfl_sync(0xffffffffu, (smem_u32(dsmem) & 0x3FFFFu) >> 4, 0);
    const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, y16 = (uint32_t)(2 * p.xbox_bytes) >> 4;
    const int nkk = p.nkk;
    uint32_t sc = 0, st = 0, ph = 0;
    int d = d0, h = h0;
    for (int l = l0; l < l1; ++l) {
      const bool act = (unsigned)(d + dz) < (unsigned)p.D && (lps > 1 || (unsigned)(h + dy) < (unsigned)p.H);
      if (++h == p.H) { h = 0; if (++d == p.D) d = 0; }
      if (!act) continue;
      mbar_wait(&st_full[st], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_s = s16 + st * stage16, b_s = a_s + y16;
        const uint32_t acc = sc > 0 ? 1u : 0u;
        // kk-major: consecutive MMAs go to different accumulators
        for (int i = 0; i < lps; ++i) {
          const uint32_t a_l = a_s + (uint32_t)(i * p.xrows * 8), b_l = b_s + (uint32_t)(i * p.W * 8);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kk < nkk) {
#pragma unroll
              for (int j = 0; j < 3; ++j)  // tap dx = j - 1 = the X line from row j on (128 B per row = 8 x 16 B)
                umma_f16(tmem_u + (uint32_t)j * 128u, wr_desc64(hi, (a_l + (uint32_t)((kk * 16 + j) * 8)) | lbo_a),
                         wr_desc64(hi, (b_l + (uint32_t)(kk * 16 * 8)) | lbo_b), idesc, (kk || i) ? 1u : acc);
            }
          }
        }
        umma_commit(&st_empty[st]);
      }
      __syncwarp();
      ++sc;
      if (++st == (uint32_t)p.stages) { st = 0; ph ^= 1u; }
    }
    if (elect_one()) {
      s_any = sc > 0;               // written (and fenced) long before the asynchronous arrive below completes
      __threadfence_block();
      umma_commit(&acc_full);       // with no MMA issued the commit arrives at once (the epilogue then skips)
    }
    __syncwarp();
  } else {
    // ===== epilogue: TMEM -> fp32 atomics into dW[widx][co][ci] =====
    const int q = warp & 3;
    const int ci = cib * 128 + q * 32 + lane;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    if (*(volatile int*)&s_any) {
      for (int j = 0; j < 3; ++j) {
        const int widx = p.lut[((dz + 1) * 3 + (dy + 1)) * 3 + j];
        if (widx < 0) continue;
        float* dst0 = p.dw + ((long long)widx * p.Cout + cob * 128) * p.Cin + ci;
        for (int c16 = 0; c16 < 8; ++c16) {
          uint32_t r[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 128 + c16 * 16), r);
          float* dst = dst0 + (long long)(c16 * 16) * p.Cin;
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

static inline int wr_align1k(long long v) { return (int)(((v + 1023) / 1024) * 1024); }

// Returns MTB200_ERR_UNSUPPORTED when the problem is outside this kernel's envelope (caller falls back).
int wgrad_rows_umma(const mtb200_wgrad_params& p, cudaStream_t s) {
  static const int on = [] { const char* e = getenv("MTB200_WGRAD_ROWS"); return e ? atoi(e) : 1; }();
  if (!on) return MTB200_ERR_UNSUPPORTED;
  if (p.ngroups != 1 || p.xform || p.ntaps != 27) return MTB200_ERR_UNSUPPORTED;
  for (int k = 0; k < 3; ++k)
    if (p.is[k] != 1 || p.os[k] != 1 || p.group_ooff[0][k] != 0) return MTB200_ERR_UNSUPPORTED;
  if (p.Do != p.Di || p.Ho != p.Hi || p.Wo != p.Wi || p.Dof != p.Do || p.Hof != p.Ho || p.Wof != p.Wo)
    return MTB200_ERR_UNSUPPORTED;
  if (p.Cin % 128 || p.Cout % 128) return MTB200_ERR_UNSUPPORTED;
  if (p.Wo != 16 && p.Wo != 32 && p.Wo != 48 && p.Wo != 64) return MTB200_ERR_UNSUPPORTED;

  static thread_local WgradRowsParams q;
  memset(&q, 0, sizeof(q));
  for (int i = 0; i < 27; ++i) q.lut[i] = -1;
  for (int t = 0; t < p.ntaps; ++t) {
    for (int k = 0; k < 3; ++k)
      if (p.tap_off[t][k] < -1 || p.tap_off[t][k] > 1) return MTB200_ERR_UNSUPPORTED;
    int& e = q.lut[((p.tap_off[t][0] + 1) * 3 + (p.tap_off[t][1] + 1)) * 3 + (p.tap_off[t][2] + 1)];
    if (e >= 0) return MTB200_ERR_UNSUPPORTED;
    e = p.tap_widx[t];
  }
  q.nkk = p.Wo / 16;
  const int xrows = p.Wo + 2;
  q.lps = (p.Wo == 16 && p.Ho % 2 == 0) ? 2 : 1;
  q.xrows = xrows;
  q.xbox_bytes = wr_align1k((long long)(q.lps * xrows + 8) * 128);  // + slack for the shifted views' last atom
  q.ybox_bytes = wr_align1k((long long)q.lps * p.Wo * 128);
  q.stage_bytes = 2 * q.xbox_bytes + 2 * q.ybox_bytes;
  q.stage_tx = q.lps * (2 * xrows * 128 + 2 * p.Wo * 128);
  q.stages = min(WR_MAX_STAGES, (200 * 1024) / q.stage_bytes);
  if (q.stages < 3) return MTB200_ERR_UNSUPPORTED;
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.Di, (cuuint64_t)p.B};
    cuuint64_t strides[4] = {(cuuint64_t)p.in_ldc * 2, (cuuint64_t)p.Wi * p.in_ldc * 2,
                             (cuuint64_t)p.Hi * p.Wi * p.in_ldc * 2, (cuuint64_t)p.Di * p.Hi * p.Wi * p.in_ldc * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)xrows, (cuuint32_t)q.lps, 1, 1};
    if (!umma_encode_map(&q.x_map, p.dtype, 5, (uint8_t*)p.x + (size_t)p.in_coff * 2, dims, strides, box, 128))
      return MTB200_ERR_CUDA;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.Cout, (cuuint64_t)p.Wof, (cuuint64_t)p.Hof, (cuuint64_t)p.Dof, (cuuint64_t)p.B};
    cuuint64_t strides[4] = {(cuuint64_t)p.out_ldc * 2, (cuuint64_t)p.Wof * p.out_ldc * 2,
                             (cuuint64_t)p.Hof * p.Wof * p.out_ldc * 2,
                             (cuuint64_t)p.Dof * p.Hof * p.Wof * p.out_ldc * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)p.Wo, (cuuint32_t)q.lps, 1, 1};
    if (!umma_encode_map(&q.dy_map, p.dtype, 5, (uint8_t*)p.dy + (size_t)p.out_coff * 2, dims, strides, box, 128))
      return MTB200_ERR_CUDA;
  }
  q.dw = p.dw;
  q.B = p.B; q.D = p.Do; q.H = p.Ho / q.lps; q.W = p.Wo;
  q.Cin = p.Cin; q.Cout = p.Cout;
  q.nci = p.Cin / 128; q.nco = p.Cout / 128;
  q.is_f16 = p.dtype == MTB200_F16;
  if ((long long)p.B * p.Do * p.Ho >= (1LL << 30)) return MTB200_ERR_UNSUPPORTED;
  q.lines = p.B * p.Do * (p.Ho / q.lps);
  const int per_slot_ctas = 9 * q.nci * q.nco;
  int nslots = max(1, num_sms() / per_slot_ctas);
  if (nslots > q.lines) nslots = q.lines;
  q.nslots = nslots;
  q.lines_per_slot = (q.lines + nslots - 1) / nslots;
  const int smem = q.stages * q.stage_bytes + 1024;
  dim3 grid((unsigned)(per_slot_ctas * nslots), 1, 1);
  cudaError_t e = cudaFuncSetAttribute(wgrad_rows_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { set_error("wgrad_rows: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  launch_pdl(wgrad_rows_umma_kernel, dim3(grid), dim3(WR_THREADS), (size_t)(smem), s, q);
  return check_launch("wgrad_rows_umma");
}

}  // namespace mtb
