// tcgen05 (UMMA) + TMA implementation of the tap-table convolution for 16-bit activations (bf16 / fp16), sm_100a.
//
// Implicit GEMM, im2col-free:  D[128 voxels x BN couts] (fp32, TMEM) = sum over (tap, channel chunk) A_tap * W_tap^T
//   * A_tap  : [128 rows = a (bd,bh,bw) brick of output voxels] x [KC channels], K-major, fetched by ONE 5-D TMA box
//              per (tap, chunk) at the tap-shifted coordinates; out-of-volume reads are zero-filled by TMA (= the conv's
//              zero padding).  Strided problems use one tensor map per input parity class (base shifted by the parity,
//              strides doubled), so no element-stride gathers are needed.
//   * W_tap  : [BN couts] x [KC channels] K-major slice of the packed weights [tap][Cout][Cin], 3-D TMA box.
//   * both operands land in shared memory in the canonical 128B/64B/32B-swizzled K-major layout (swizzle span = KC*2 B)
//     and are consumed by tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) issued by one thread.
//   * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
//     (TMEM -> registers -> +bias -> round -> global; per-(b, cout) sum / sum-of-squares for InstanceNorm).
//   * one output tile per CTA, two CTAs per SM (<= 256 TMEM columns and <= ~100 KB smem each) so that one CTA's
//     epilogue overlaps the other's main loop.
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"

namespace mtb {

// ---------------------------------------------------------------------------------------------------------------------
// driver entry point for tensor-map encoding (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------------------------------
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
  }
  return fn;
}

int conv_halo_umma(const mtb200_conv_params& p, cudaStream_t s);  // conv_halo.cu
int conv_line_umma(const mtb200_conv_params& p, cudaStream_t s);  // conv_line.cu
int wgrad_line_umma(const mtb200_wgrad_params& p, cudaStream_t s);  // wgrad_line.cu
int wgrad_line_s2_umma(const mtb200_wgrad_params& p, cudaStream_t s);  // wgrad_line_s2.cu
int wgrad_rows_umma(const mtb200_wgrad_params& p, cudaStream_t s);     // wgrad_rows.cu
int conv_pw_umma(const mtb200_conv_params& p, cudaStream_t s);  // conv_pw.cu
int conv_gm_umma(const mtb200_conv_params& p, cudaStream_t s);  // conv_gm.cu

int umma_available() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0;
    cached = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess && major == 10 &&
        get_encode_fn() != nullptr)
      cached = 1;
    cudaGetLastError();
  }
  return cached;
}

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers: umma.cuh
// ---------------------------------------------------------------------------------------------------------------------
using namespace um;
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// shared-memory matrix descriptor, K-major, swizzle span == row pitch (KC * 2 bytes), atoms of 8 rows
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t row_bytes) {
  const uint32_t sbo = 8u * row_bytes;  // byte distance between 8-row groups
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);  // SWIZZLE_128B / 64B / 32B
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address            bits [0,14)
  d |= (uint64_t)0 << 16;                              // leading byte offset      bits [16,30) (unused: 1 atom on K)
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;         // stride byte offset       bits [32,46)
  d |= 1ull << 46;                                     // descriptor version (sm_100)
  d |= layout << 61;                                   // swizzle mode             bits [61,64)
  return d;
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------------------
constexpr int UM_MAX_STAGES = 12;
constexpr int UM_THREADS = 192;       // weight-gradient kernel: producer, MMA, 4 epilogue warps
constexpr int UMC_THREADS = 192;      // conv kernel: producer, MMA, 4 epilogue warps (two CTAs per SM when N <= 128)

struct UmmaConvParams {
  CUtensorMap a_maps[8];
  CUtensorMap w_map;
  void* out;
  const float* bias;
  double* stats;
  int B, Do, Ho, Wo;
  int bd, bh, bw, tiles_d, tiles_h, tiles_w;
  int Dof, Hof, Wof, out_ldc, out_coff, Cout;
  int os[3];
  int group_tap_begin[MTB200_MAX_GROUPS + 1];
  int group_ooff[MTB200_MAX_GROUPS][3];
  int tap_map[MTB200_MAX_TAPS];
  int tap_coff[MTB200_MAX_TAPS][3];
  int tap_widx[MTB200_MAX_TAPS];
  int nkc, KC, BN, stages, tmem_cols;
  int ny, ncombo, ctas_per_combo;   // N tiles, (N tile, group, split) combinations, persistent CTAs per combination
  int ngroups_k;                    // tap groups
  long long ntiles;                 // spatial tiles (all batches)
  int accumulate, is_f16;
  // split over taps (deep levels: a handful of tiles would leave most SMs idle): combination = (N tile, group, split),
  // split k accumulates taps [k * taps/nsplit, ...) and stores its fp32 partial tile into slab k of `scratch`
  // ([nsplit][voxels][Cout]); split_reduce_kernel adds the slabs in a fixed order, + bias, rounds, stores, statistics
  int nsplit;
  float* scratch;
  long long mtot;                   // output voxels (all batches)
};

// Persistent: CTA c owns the (N tile, tap group) combination c % ncombo and walks the spatial tiles c / ncombo,
// c / ncombo + ctas_per_combo, ...  Two TMEM accumulators alternate, so the epilogue warps drain tile k while the
// TMA / MMA warps are already on tile k+1; barrier setup, TMEM allocation and descriptor fetch are paid once per CTA.
// Two CTAs share an SM when the N tile is <= 128 (2 x 2 x BN <= 512 TMEM columns): their MMA-issue loops interleave,
// which hides the per-stage barrier round trip of the single issuing thread.
//
// PAIR (r2, `MTB200_TAPS_PAIR`): the CTAs of a 2-cluster (two SMs of one TPC) work on two consecutive spatial tiles with ONE
// tcgen05.mma.cta_group::2 (M = 256): each CTA fetches its own 128-voxel activation box and HALF of the weight tile's N
// rows, so the weight stream through L2 -> SM (the bound of this kernel) is halved and N tiles up to 256 fit.  Only the
// rank-0 CTA issues MMAs and owns the `full` barriers (both CTAs' TMA bytes are counted there); tcgen05.commit multicasts
// the "slot free" / "accumulator complete" arrivals to both CTAs, the epilogue warps of both arrive on rank 0's
// `acc_empty`.
template <typename T, bool PAIR>
__global__ void __launch_bounds__(UMC_THREADS, 2) conv_taps_umma_kernel(const __grid_constant__ UmmaConvParams p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t full_bar[UM_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[UM_MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  // statistics: one slot per epilogue warp (TMEM lane quarter), summed in a fixed order at the flush -- no float atomics,
  // so the forward pass is reproducible run to run (DESIGN.md "Run-to-run variation")
  __shared__ float s_sum[4][256], s_sq[4][256], s_bias[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t row_bytes = p.KC * 2;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;  // cluster dims (2, 1, 1): rank == blockIdx.x & 1
  const uint32_t b_rows = PAIR ? (uint32_t)p.BN / 2u : (uint32_t)p.BN;  // weight rows this CTA fetches
  const uint32_t a_bytes = 128u * row_bytes, b_bytes = b_rows * row_bytes;
  const uint32_t stage_bytes = ((a_bytes + b_bytes + 1023u) / 1024u) * 1024u;

  // a "unit" is what one MMA chain covers: one spatial tile, or the pair's two consecutive tiles (2u, 2u + 1)
  const int unit_cta = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int combo = unit_cta % p.ncombo;
  const int nb = combo % p.ny, g = (combo / p.ny) % p.ngroups_k, ks = combo / (p.ny * p.ngroups_k);
  const long long tile0 = unit_cta / p.ncombo;
  const long long nunits = PAIR ? (p.ntiles + 1) / 2 : p.ntiles;
  const int n0 = nb * p.BN;
  const int taps_per = (p.group_tap_begin[g + 1] - p.group_tap_begin[g]) / p.nsplit;  // host: divisible
  const int tap_begin = p.group_tap_begin[g] + ks * taps_per, tap_end = tap_begin + taps_per;
  const int niter = (tap_end - tap_begin) * p.nkc;
  const long long tiles_per_b = (long long)p.tiles_d * p.tiles_h * p.tiles_w;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], PAIR ? 8 : 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += UMC_THREADS) {
#pragma unroll
    for (int w = 0; w < 4; ++w) { s_sum[w][i] = 0.f; s_sq[w][i] = 0.f; }
    s_bias[i] = (p.bias && i < p.BN) ? p.bias[n0 + i] : 0.f;
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(&tmem_slot, (uint32_t)p.tmem_cols);  // the same warp of both CTAs: same columns in both TMEMs
    else tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (warp-uniform loop, one elected lane issues) =====
    if (niter > 0) {
      uint32_t gi = 0;  // global pipeline iteration of this CTA
      for (long long u = tile0; u < nunits; u += p.ctas_per_combo) {
        // ntiles < 2^31 (host check): 32-bit divisions, a fifth of the 64-bit ones' instructions.  The odd pair tile past
        // the end decodes to b == B: its box is out of bounds on the batch axis and arrives as zeros
        uint32_t t = PAIR ? (uint32_t)(2 * u) + rank : (uint32_t)u;
        const int tw = (int)(t % (uint32_t)p.tiles_w); t /= (uint32_t)p.tiles_w;
        const int th = (int)(t % (uint32_t)p.tiles_h); t /= (uint32_t)p.tiles_h;
        const int td = (int)(t % (uint32_t)p.tiles_d);
        const int b = (int)(t / (uint32_t)p.tiles_d);
        const int d0 = td * p.bd, h0 = th * p.bh, w0 = tw * p.bw;
        int tp = tap_begin, kc = 0;
        for (int it = 0; it < niter; ++it, ++gi) {
          const uint32_t stage = gi % (uint32_t)p.stages;
          mbar_wait(&empty_bar[stage], ((gi / (uint32_t)p.stages) & 1u) ^ 1u);
          if (elect_one()) {
            uint8_t* sa = dsmem + (size_t)stage * stage_bytes;
            if (PAIR) {
              if (rank == 0) mbar_expect_tx(&full_bar[stage], 2u * (a_bytes + b_bytes));  // both CTAs' boxes
              tma_load_5d_pair(sa, &p.a_maps[p.tap_map[tp]], &full_bar[stage], kc * p.KC, w0 + p.tap_coff[tp][2],
                               h0 + p.tap_coff[tp][1], d0 + p.tap_coff[tp][0], b);
              tma_load_3d_pair(sa + a_bytes, &p.w_map, &full_bar[stage], kc * p.KC, n0 + (int)(rank * b_rows),
                               p.tap_widx[tp]);
            } else {
              mbar_expect_tx(&full_bar[stage], a_bytes + b_bytes);
              tma_load_5d(sa, &p.a_maps[p.tap_map[tp]], &full_bar[stage], kc * p.KC, w0 + p.tap_coff[tp][2],
                          h0 + p.tap_coff[tp][1], d0 + p.tap_coff[tp][0], b);
              tma_load_3d(sa + a_bytes, &p.w_map, &full_bar[stage], kc * p.KC, n0, p.tap_widx[tp]);
            }
          }
          __syncwarp();
          if (++kc == p.nkc) { kc = 0; ++tp; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues (descriptors stay in uniform registers) =====
    if (niter > 0 && rank == 0) {
      // instruction descriptor: D=f32, A/B = bf16 or f16, both K-major, N, M=128 (256 across the pair)
      const uint32_t fmt = p.is_f16 ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) |
                             (((PAIR ? 256u : 128u) >> 4) << 24);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t s16 = __shfl_sync(0xffffffffu, (smem_u32(dsmem) & 0x3FFFFu) >> 4, 0);
      const uint32_t layout = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);
      const uint32_t hi = (((8u * row_bytes) >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
      const uint32_t stage16 = stage_bytes >> 4, a16 = a_bytes >> 4;
      const int ksteps = p.KC / 16;
      uint32_t gi = 0, k = 0;
      for (long long u = tile0; u < nunits; u += p.ctas_per_combo, ++k) {
        const uint32_t buf = k & 1u;
        mbar_wait(&acc_empty[buf], ((k >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t dcol = tmem_u + buf * (uint32_t)p.BN;
        for (int it = 0; it < niter; ++it, ++gi) {
          const uint32_t stage = gi % (uint32_t)p.stages;
          mbar_wait(&full_bar[stage], (gi / (uint32_t)p.stages) & 1u);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = s16 + stage * stage16;
            const uint32_t sb = sa + a16;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t ad = ((uint64_t)hi << 32) | (uint64_t)(sa + 2u * ks), bd = ((uint64_t)hi << 32) | (uint64_t)(sb + 2u * ks);
              if (PAIR) umma_f16_pair(dcol, ad, bd, idesc, (it > 0 || ks > 0) ? 1u : 0u);
              else umma_f16(dcol, ad, bd, idesc, (it > 0 || ks > 0) ? 1u : 0u);
            }
            if (PAIR) {
              umma_commit_pair(&empty_bar[stage]);  // frees the slot in BOTH CTAs
              if (it == niter - 1) umma_commit_pair(&acc_full[buf]);
            } else {
              umma_commit(&empty_bar[stage]);  // frees the smem slot once the MMAs above have read it
              if (it == niter - 1) umma_commit(&acc_full[buf]);  // accumulator complete
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue: warps 2..5; warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
    const int q = warp & 3;
    const int row = q * 32 + lane;  // row of the tile == TMEM lane
    const int rw = row % p.bw, rh = (row / p.bw) % p.bh, rd = row / (p.bw * p.bh);
    T* out = reinterpret_cast<T*>(p.out);
    const bool want_stats = p.stats != nullptr;
    uint32_t k = 0;
    for (long long u = tile0; u < nunits; u += p.ctas_per_combo, ++k) {
      const long long tile = PAIR ? 2 * u + rank : u;
      uint32_t t = (uint32_t)tile;
      const int tw = (int)(t % (uint32_t)p.tiles_w); t /= (uint32_t)p.tiles_w;
      const int th = (int)(t % (uint32_t)p.tiles_h); t /= (uint32_t)p.tiles_h;
      const int td = (int)(t % (uint32_t)p.tiles_d);
      const int b = (int)(t / (uint32_t)p.tiles_d);
      const int od = td * p.bd + rd, oh = th * p.bh + rh, ow = tw * p.bw + rw;
      const bool valid = od < p.Do && oh < p.Ho && ow < p.Wo && (!PAIR || tile < p.ntiles);
      const long long ovox = (((long long)b * p.Dof + (od * p.os[0] + p.group_ooff[g][0])) * p.Hof +
                              (oh * p.os[1] + p.group_ooff[g][1])) * p.Wof + (ow * p.os[2] + p.group_ooff[g][2]);
      T* orow = out + ovox * p.out_ldc + p.out_coff + n0;
      const uint32_t buf = k & 1u;
      const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)p.BN;
      if (niter > 0) {
        mbar_wait(&acc_full[buf], (k >> 1) & 1u);
        tc_fence_after();
      }
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        float v[16];
        if (niter > 0) {
          uint32_t r[16];
          tmem_ld16(tcol + (uint32_t)c0, r);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + s_bias[c0 + j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = s_bias[c0 + j];
        }
        if (p.nsplit > 1) {  // fp32 partial tile of this tap split (no bias / statistics: split_reduce_kernel)
          if (valid) {
            float4* sc = reinterpret_cast<float4*>(p.scratch + ((size_t)ks * p.mtot + ovox) * p.Cout + n0 + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          continue;
        }
        if (valid) {
          if (p.accumulate) {
            float o[8];
            load8<T>(orow + c0, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += o[j];
            load8<T>(orow + c0 + 8, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[8 + j] += o[j];
          }
          float lo[8], hi8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi8[j] = v[8 + j]; }
          store8<T>(orow + c0, lo);
          store8<T>(orow + c0 + 8, hi8);
        }
        if (want_stats) {
          // column sums over the 32 rows of this warp: transposing butterfly (16 shuffles per statistic)
          float sv[16], ss[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = valid ? Traits<T>::round(v[j]) : 0.f;
            sv[j] = x;
            ss[j] = x * x;
          }
          warp_colsum16(sv, lane);
          warp_colsum16(ss, lane);
          if ((lane & 1) == 0) {
            const int col = colsum16_column(lane);
            s_sum[q][c0 + col] += sv[0];  // (q, column) has exactly one owner lane
            s_sq[q][c0 + col] += ss[0];
          }
        }
      }
      if (niter > 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_leader(&acc_empty[buf]);
          else mbar_arrive(&acc_empty[buf]);
        }
      }
      if (want_stats) {
        // flush the per-(b, channel) partials when this CTA moves on to another sample (or is done)
        const long long next = tile + (PAIR ? 2 : 1) * p.ctas_per_combo;
        if (next >= p.ntiles || next / tiles_per_b != tile / tiles_per_b) {
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int c = threadIdx.x - 64; c < p.BN; c += 128) {
            const float su = ((s_sum[0][c] + s_sum[1][c]) + s_sum[2][c]) + s_sum[3][c];
            const float sq = ((s_sq[0][c] + s_sq[1][c]) + s_sq[2][c]) + s_sq[3][c];
            if (su != 0.f || sq != 0.f) {
              double* st = p.stats + ((long long)b * p.Cout + n0 + c) * 2;
              atomicAdd(st, (double)su);
              atomicAdd(st + 1, (double)sq);
#pragma unroll
              for (int w = 0; w < 4; ++w) { s_sum[w][c] = 0.f; s_sq[w][c] = 0.f; }
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other may still signal it / read its tiles
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// Second half of a tap-split launch: out[v][c] = round(bias[c] + sum_k scratch[k][v][c] (+ out[v][c])), slabs added in
// the fixed order k = 0, 1, ... (run-to-run reproducible), InstanceNorm statistics of the rounded values.  Block =
// (Cout / 8 channel groups) x rows; grid = (chunks of a sample's voxels, B).
template <typename T>
__global__ void __launch_bounds__(256) split_reduce_kernel(const float* __restrict__ scratch, int nsplit, long long mtot,
                                                           int vox_per_b, int Cout, T* __restrict__ out, int out_ldc,
                                                           int out_coff, const float* __restrict__ bias,
                                                           double* __restrict__ stats, int accumulate, int chunk) {
  pdl_wait();
  __shared__ float s_red[256][17];
  const int ncg = Cout >> 3;
  const int rows = 256 / ncg;
  const int cg = threadIdx.x % ncg, r = threadIdx.x / ncg;
  const int b = blockIdx.y;
  const int v0 = blockIdx.x * chunk, v1 = min(vox_per_b, v0 + chunk);
  float bs[8], su[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { bs[j] = bias ? bias[cg * 8 + j] : 0.f; su[j] = 0.f; sq[j] = 0.f; }
  if (r < rows) {
    for (int v = v0 + r; v < v1; v += rows) {
      const long long gv = (long long)b * vox_per_b + v;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int k = 0; k < nsplit; ++k) {
        float x[8];
        load8<float>(scratch + ((size_t)k * mtot + gv) * Cout + cg * 8, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += x[j];
      }
      T* o = out + gv * out_ldc + out_coff + cg * 8;
      float old[8];
      if (accumulate) load8<T>(o, old);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += bs[j] + (accumulate ? old[j] : 0.f);
      store8<T>(o, acc);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x = Traits<T>::round(acc[j]);
        su[j] += x;
        sq[j] = fmaf(x, x, sq[j]);
      }
    }
  }
  if (stats) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_red[threadIdx.x][j] = su[j]; s_red[threadIdx.x][8 + j] = sq[j]; }
    __syncthreads();
    // one thread per (channel, statistic): fixed summation order over the rows
    for (int i = threadIdx.x; i < Cout * 2; i += 256) {
      const int c = i >> 1, which = i & 1;
      float t = 0.f;
      for (int rr = 0; rr < rows; ++rr) t += s_red[rr * ncg + (c >> 3)][which * 8 + (c & 7)];
      if (t != 0.f) atomicAdd(stats + ((long long)b * Cout + c) * 2 + which, (double)t);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

static bool encode_map(EncodeTiledFn enc, CUtensorMap* m, CUtensorMapDataType dt, int rank, void* base,
                       const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box, int row_bytes) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                    : (row_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_32B));
  CUresult r = enc(m, dt, (cuuint32_t)rank, base, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu box %u %u %u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0], box[1], box[2]);
    return false;
  }
  return true;
}

static bool g_pair_refused = false;  // set when a cluster launch of the PAIR instantiation was refused

// second half of a tap-split launch (see UmmaConvParams::nsplit); frees the slabs in stream order
static int split_reduce(const mtb200_conv_params& p, int nsplit, float* scratch, int status, cudaStream_t s) {
  if (status == MTB200_OK) {
    const long long M = (long long)p.B * p.Do * p.Ho * p.Wo;
    const int vox_per_b = p.Do * p.Ho * p.Wo;
    const int rows = 256 / (p.Cout / 8);
    // ~4 blocks per SM over all samples; a block walks `chunk` voxels, `rows` at a time
    int chunks = max(1, min((vox_per_b + rows - 1) / rows, (4 * num_sms() + p.B - 1) / p.B));
    const int chunk = (vox_per_b + chunks - 1) / chunks;
    chunks = (vox_per_b + chunk - 1) / chunk;
    dim3 grid((unsigned)chunks, (unsigned)p.B, 1);
    if (p.dtype == MTB200_BF16)
      launch_pdl(split_reduce_kernel<__nv_bfloat16>, grid, dim3(256), (size_t)0, s, (const float*)scratch, nsplit, M, vox_per_b,
                 p.Cout, (__nv_bfloat16*)p.out, p.out_ldc, p.out_coff, p.bias, p.stats, p.accumulate, chunk);
    else
      launch_pdl(split_reduce_kernel<__half>, grid, dim3(256), (size_t)0, s, (const float*)scratch, nsplit, M, vox_per_b,
                 p.Cout, (__half*)p.out, p.out_ldc, p.out_coff, p.bias, p.stats, p.accumulate, chunk);
    status = check_launch("split_reduce");
  }
  cudaFreeAsync(scratch, s);
  return status;
}

int conv_taps_umma(const mtb200_conv_params& p, cudaStream_t s) {
  if (!umma_available()) { set_error("conv_taps(umma): no sm_100 device / driver entry point"); return MTB200_ERR_UNSUPPORTED; }
  if (p.dtype != MTB200_BF16 && p.dtype != MTB200_F16) { set_error("conv_taps(umma): 16-bit activations only"); return MTB200_ERR_UNSUPPORTED; }
  if (p.wdtype != p.dtype) { set_error("conv_taps(umma): weights must have the activation dtype"); return MTB200_ERR_UNSUPPORTED; }
  if (p.xform) { set_error("conv_taps(umma): on-load transform not supported (materialise the input)"); return MTB200_ERR_UNSUPPORTED; }
  if (p.Cin % 16 || p.Cout % 16 || p.in_ldc % 8 || p.in_coff % 8 || p.out_ldc % 8 || p.out_coff % 8) {
    set_error("conv_taps(umma): channel alignment (Cin %d Cout %d)", p.Cin, p.Cout);
    return MTB200_ERR_UNSUPPORTED;
  }
  for (int k = 0; k < 3; ++k)
    if (p.is[k] < 1 || p.is[k] > 2) { set_error("conv_taps(umma): input stride %d", p.is[k]); return MTB200_ERR_UNSUPPORTED; }
  const long long M = (long long)p.B * p.Do * p.Ho * p.Wo;
  if (M == 0) return MTB200_OK;
  // impl 3 = per-tap kernel only, 4 = plane-streaming only, 5 = line-streaming (dy merged into N) only, 6 = pointwise
  // streaming only; auto = what measured fastest on the B200 (tools/conv_bench.py, profiles/): the flat streaming GEMM
  // for 1x1x1 layers, line-streaming for the wide-W narrow-channel layers, plane-streaming for other narrow layers whose
  // 27 weight tiles stay resident, per-tap otherwise.
  if (p.impl == 0 || p.impl == 2 || p.impl == 6) {
    const int r = conv_pw_umma(p, s);
    if (r != MTB200_ERR_UNSUPPORTED) return r;
    if (p.impl == 6) { set_error("conv_taps(umma): problem outside the pointwise kernel's envelope"); return r; }
  }
  if (p.impl == 0 || p.impl == 2 || p.impl == 7) {  // 7 = group-merged lattice kernel only
    const int r = conv_gm_umma(p, s);
    if (r != MTB200_ERR_UNSUPPORTED) return r;
    if (p.impl == 7) { set_error("conv_taps(umma): problem outside the group-merged kernel's envelope"); return r; }
  }
  if (p.impl == 5 || (p.impl != 3 && p.impl != 4 && p.impl != 6)) {
    const int r = conv_line_umma(p, s);
    if (r != MTB200_ERR_UNSUPPORTED) return r;
    if (p.impl == 5) { set_error("conv_taps(umma): problem outside the line-streaming kernel's envelope"); return r; }
  }
  if (p.impl == 4 || (p.impl != 3 && p.Cin <= 32 && p.Cout <= 32)) {
    const int r = conv_halo_umma(p, s);
    if (r != MTB200_ERR_UNSUPPORTED) return r;
    if (p.impl == 4) { set_error("conv_taps(umma): problem outside the plane-streaming kernel's envelope"); return r; }
  }
  EncodeTiledFn enc = get_encode_fn();

  static thread_local UmmaConvParams q;  // large: keep off the stack; one instance per host thread (re-entrant per thread)
  memset(&q, 0, sizeof(q));
  q.KC = (p.Cin % 64 == 0) ? 64 : ((p.Cin % 32 == 0) ? 32 : 16);
  q.nkc = p.Cin / q.KC;
  // N tile: whole Cout if <= 128, else the largest multiple-of-16 divisor <= 160
  q.BN = p.Cout;
  if (q.BN > 128) {
    q.BN = 0;
    for (int c = 160; c >= 16; c -= 16)
      if (p.Cout % c == 0) { q.BN = c; break; }
  }
  // CTA pairs (cta_group::2, M = 256): one tap group with taps, >= 2 tiles; N tile = whole Cout up to 256 (each CTA of
  // the pair fetches half of its rows).  MTB200_TAPS_PAIR=0 switches the pairs off (5 - 20 % slower on every
  // layer of the stack, profiles/r2w_pair.txt)
  bool pair = false;
  {
    static const int mode = [] { const char* e = getenv("MTB200_TAPS_PAIR"); return e ? atoi(e) : -1; }();
    // (64-byte rows, i.e. the strided 32 -> 64 layer at full resolution, measured 30 % SLOWER in pairs)
    bool eligible = (p.ngroups == 1 || mode == 2) && p.Cout % 32 == 0 && q.KC == 64;  // mode 2: multi-group problems too
    for (int g = 0; g < p.ngroups; ++g) eligible = eligible && p.group_tap_begin[g + 1] > p.group_tap_begin[g];
    int bn2 = p.Cout;
    if (bn2 > 256) {
      bn2 = 0;
      for (int c = 256; c >= 32; c -= 32)
        if (p.Cout % c == 0) { bn2 = c; break; }
    }
    eligible = eligible && bn2 >= 32;
    if (mode != 0 && !g_pair_refused) pair = eligible;
    if (pair) q.BN = bn2;
  }
  q.tmem_cols = 32;
  while (q.tmem_cols < 2 * q.BN) q.tmem_cols *= 2;  // two accumulators (tile k drains while tile k+1 accumulates)
  // brick: powers of two with product 128 minimising the number of tiles (ties: widest in w)
  long long best = -1;
  for (int bw = 128; bw >= 1; bw >>= 1)
    for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
      const int bd = 128 / (bw * bh);
      const long long nt = (long long)((p.Do + bd - 1) / bd) * ((p.Ho + bh - 1) / bh) * ((p.Wo + bw - 1) / bw);
      if (best < 0 || nt < best) { best = nt; q.bd = bd; q.bh = bh; q.bw = bw; }
    }
  q.tiles_d = (p.Do + q.bd - 1) / q.bd; q.tiles_h = (p.Ho + q.bh - 1) / q.bh; q.tiles_w = (p.Wo + q.bw - 1) / q.bw;
  const long long ntiles = (long long)p.B * q.tiles_d * q.tiles_h * q.tiles_w;
  MTB_REQUIRE(ntiles < (1LL << 31), "conv_taps(umma): too many tiles");

  const int row_bytes = q.KC * 2;
  const CUtensorMapDataType dt = p.dtype == MTB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  // one activation map per parity class used by the taps
  int parity_slot[8];
  for (int i = 0; i < 8; ++i) parity_slot[i] = -1;
  int nmaps = 0;
  const int Dims[3] = {p.Di, p.Hi, p.Wi};
  for (int t = 0; t < p.ntaps; ++t) {
    int par[3], coord[3];
    for (int k = 0; k < 3; ++k) {
      const int off = p.tap_off[t][k];
      par[k] = ((off % p.is[k]) + p.is[k]) % p.is[k];
      coord[k] = floor_div(off - par[k], p.is[k]);
    }
    const int code = par[0] * 4 + par[1] * 2 + par[2];
    if (parity_slot[code] < 0) {
      // extents of the parity sub-lattice; an empty sub-lattice (possible only for degenerate sizes) is unsupported
      cuuint64_t dims[5], strides[4];
      cuuint32_t box[5] = {(cuuint32_t)q.KC, (cuuint32_t)q.bw, (cuuint32_t)q.bh, (cuuint32_t)q.bd, 1};
      long long ext[3];
      for (int k = 0; k < 3; ++k) {
        ext[k] = (Dims[k] - par[k] + p.is[k] - 1) / p.is[k];
        if (ext[k] < 1) { set_error("conv_taps(umma): empty parity lattice"); return MTB200_ERR_UNSUPPORTED; }
      }
      dims[0] = p.Cin; dims[1] = ext[2]; dims[2] = ext[1]; dims[3] = ext[0]; dims[4] = p.B;
      const long long e = 2;  // bytes per element
      strides[0] = (cuuint64_t)p.in_ldc * e * p.is[2];
      strides[1] = (cuuint64_t)p.Wi * p.in_ldc * e * p.is[1];
      strides[2] = (cuuint64_t)p.Hi * p.Wi * p.in_ldc * e * p.is[0];
      strides[3] = (cuuint64_t)p.Di * p.Hi * p.Wi * p.in_ldc * e;
      uint8_t* base = (uint8_t*)p.in + ((((long long)par[0] * p.Hi + par[1]) * p.Wi + par[2]) * p.in_ldc + p.in_coff) * e;
      if (!encode_map(enc, &q.a_maps[nmaps], dt, 5, base, dims, strides, box, row_bytes)) return MTB200_ERR_CUDA;
      parity_slot[code] = nmaps++;
    }
    q.tap_map[t] = parity_slot[code];
    for (int k = 0; k < 3; ++k) q.tap_coff[t][k] = coord[k];
    q.tap_widx[t] = p.tap_widx[t];
  }
  {
    int n_widx = 0;
    for (int t = 0; t < p.ntaps; ++t) n_widx = max(n_widx, p.tap_widx[t] + 1);
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, (cuuint64_t)n_widx};
    cuuint64_t strides[2] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)p.Cin * p.Cout * 2};
    cuuint32_t box[3] = {(cuuint32_t)q.KC, (cuuint32_t)(pair ? q.BN / 2 : q.BN), 1};
    if (!encode_map(enc, &q.w_map, dt, 3, (void*)p.w, dims, strides, box, row_bytes)) return MTB200_ERR_CUDA;
  }
  q.out = p.out; q.bias = p.bias; q.stats = p.stats;
  q.B = p.B; q.Do = p.Do; q.Ho = p.Ho; q.Wo = p.Wo;
  q.Dof = p.Dof; q.Hof = p.Hof; q.Wof = p.Wof; q.out_ldc = p.out_ldc; q.out_coff = p.out_coff; q.Cout = p.Cout;
  for (int k = 0; k < 3; ++k) q.os[k] = p.os[k];
  for (int g = 0; g <= p.ngroups; ++g) q.group_tap_begin[g] = p.group_tap_begin[g];
  for (int g = 0; g < p.ngroups; ++g)
    for (int k = 0; k < 3; ++k) q.group_ooff[g][k] = p.group_ooff[g][k];
  q.accumulate = p.accumulate;
  q.is_f16 = p.dtype == MTB200_F16;

  const int stage_bytes = ((128 * row_bytes + (pair ? q.BN / 2 : q.BN) * row_bytes + 1023) / 1024) * 1024;
  const int per_sm = q.tmem_cols <= 256 ? 2 : 1;  // CTAs per SM (TMEM: 512 columns per SM)
  const int budget = per_sm == 2 ? 100 * 1024 : 200 * 1024;
  q.stages = max(2, min(UM_MAX_STAGES, budget / stage_bytes));
  int smem = q.stages * stage_bytes + 1024;
  if (per_sm == 1) smem = max(smem, 116 * 1024);
  q.ntiles = ntiles;
  q.ny = p.Cout / q.BN;
  q.ngroups_k = p.ngroups;
  q.nsplit = 1;
  q.mtot = M;
  const int slots = num_sms() * per_sm;
  // Tap split (MTB200_TAPS_SPLIT=0: off): the deep levels have 8 - 30 tiles, so most SMs would idle while a few CTAs walk
  // 27 taps x Cin / 64 chunks each.  Split the taps into nsplit equal runs (fp32 partial slabs + a reduce kernel) when the
  // unsplit launch would fill less than half of the machine; nsplit minimises rounds / nsplit + the reduce pass's share.
  float* scratch = nullptr;
  {
    static const int split_on = [] { const char* e = getenv("MTB200_TAPS_SPLIT"); return e ? atoi(e) : 1; }();  // > 1: forced
    const long long units = (pair ? (ntiles + 1) / 2 : ntiles) * q.ny;
    const long long uslots = pair ? slots / 2 : slots;
    bool ok = split_on && p.ngroups == 1 && units * 2 <= uslots && p.Cout % 16 == 0 && p.Cout <= 2048 &&
              p.Dof == p.Do && p.Hof == p.Ho && p.Wof == p.Wo;
    for (int k = 0; k < 3; ++k) ok = ok && p.os[k] == 1 && p.group_ooff[0][k] == 0;
    if (ok) {
      const int nt = p.group_tap_begin[1] - p.group_tap_begin[0];
      // relative cost model fitted to tools/conv_bench.py (profiles/r2w_tap_split.txt): rounds / nsplit for the main loop,
      // 0.05 per split for the slabs' traffic; unsplit = 1, a split must promise < 0.85
      double best = 0.85;
      for (int ns = 2; ns <= nt; ++ns) {
        if (nt % ns) continue;
        const long long rounds = (units * ns + uslots - 1) / uslots;
        const double cost = (double)rounds / ns + 0.05 * ns;
        if (cost < best) { best = cost; q.nsplit = ns; }
      }
      if (split_on > 1 && nt % split_on == 0) q.nsplit = split_on;
    }
    if (q.nsplit > 1) {
      const size_t bytes = (size_t)q.nsplit * (size_t)M * p.Cout * sizeof(float);
      // the device's default pool gives freed blocks back to the driver at the next synchronisation unless told to keep
      // them (a fresh physical allocation per call costs ~0.3 ms)
      static thread_local int pool_dev = -1;
      int dev = 0;
      cudaGetDevice(&dev);
      if (pool_dev != dev) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
          unsigned long long keep = 1ull << 30;
          cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        pool_dev = dev;
      }
      if (cudaMallocAsync(&scratch, bytes, s) != cudaSuccess) { cudaGetLastError(); q.nsplit = 1; scratch = nullptr; }
    }
    if (q.nsplit > 1) { q.scratch = scratch; q.bias = nullptr; q.stats = nullptr; q.accumulate = 0; }
  }
  q.ncombo = q.ny * p.ngroups * q.nsplit;
  if (q.ncombo > slots) { set_error("conv_taps(umma): %d (N tile, group) combinations exceed the CTA slots", q.ncombo); return MTB200_ERR_UNSUPPORTED; }
  cudaError_t e;
  if (pair) {
    // persistent PAIRS: ctas_per_combo counts pairs, the grid is twice that (cluster dims (2, 1, 1))
    const int pair_slots = slots / 2;
    if (q.ncombo > pair_slots) { set_error("conv_taps(umma): %d N tiles exceed the CTA pair slots", q.ncombo); return MTB200_ERR_UNSUPPORTED; }
    q.ctas_per_combo = (int)min((long long)(pair_slots / q.ncombo), (ntiles + 1) / 2);
    dim3 grid((unsigned)(2 * q.ctas_per_combo * q.ncombo), 1, 1);
    if (p.dtype == MTB200_BF16) {
      e = cudaFuncSetAttribute(conv_taps_umma_kernel<__nv_bfloat16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e == cudaSuccess) e = launch_pdl_cluster(conv_taps_umma_kernel<__nv_bfloat16, true>, grid, dim3(UMC_THREADS), (size_t)smem, s, 2, q);
    } else {
      e = cudaFuncSetAttribute(conv_taps_umma_kernel<__half, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e == cudaSuccess) e = launch_pdl_cluster(conv_taps_umma_kernel<__half, true>, grid, dim3(UMC_THREADS), (size_t)smem, s, 2, q);
    }
    if (e != cudaSuccess) {
      // a device / driver that refuses 2-CTA cluster launches (MIG slices, cluster scheduling disabled): nothing was
      // launched, use the single-CTA instantiation from now on
      cudaGetLastError();
      if (scratch) cudaFreeAsync(scratch, s);
      g_pair_refused = true;
      return conv_taps_umma(p, s);
    }
    const int r = check_launch("conv_taps_umma(pair)");
    return scratch ? split_reduce(p, q.nsplit, scratch, r, s) : r;
  }
  q.ctas_per_combo = (int)min((long long)(slots / q.ncombo), ntiles);
  dim3 grid((unsigned)(q.ctas_per_combo * q.ncombo), 1, 1);
  if (p.dtype == MTB200_BF16) {
    e = cudaFuncSetAttribute(conv_taps_umma_kernel<__nv_bfloat16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) launch_pdl(conv_taps_umma_kernel<__nv_bfloat16, false>, dim3(grid), dim3(UMC_THREADS), (size_t)(smem), s, q);
  } else {
    e = cudaFuncSetAttribute(conv_taps_umma_kernel<__half, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) launch_pdl(conv_taps_umma_kernel<__half, false>, dim3(grid), dim3(UMC_THREADS), (size_t)(smem), s, q);
  }
  if (e != cudaSuccess) { set_error("conv_taps(umma): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); if (scratch) cudaFreeAsync(scratch, s); return MTB200_ERR_CUDA; }
  const int r = check_launch("conv_taps_umma");
  return scratch ? split_reduce(p, q.nsplit, scratch, r, s) : r;
}

// =====================================================================================================================
// weight gradient on tcgen05:  dW[t][co][ci] += sum_v dY[v][co] * X[v + off_t][ci]
//
// GEMM with K = voxels (split over CTAs), both operands MN-major (channels are the contiguous axis):
//   A  [K = 64 voxels][M = 128 rows = (tap, ci) "row blocks"]   each row block = the tap-shifted X brick, one TMA box
//   B  [K = 64 voxels][N = BN couts]                            the dY brick
//   D_j[128 x BN] fp32 in TMEM, one accumulator per M-tile j handled by this CTA (sum of BN over tiles <= 512 columns)
// A row block is `cb` channels wide (cb = 64 / 32 / 16 = the 128B / 64B / 32B swizzle span); blocks of one M-tile sit at
// a uniform distance (= LBO) in shared memory, 8-voxel groups at SBO = 8 * cb * 2 bytes.
// Grid: x = split over voxel bricks, y = group of M-tiles (taps), z = N tile.  Epilogue: fp32 atomics into dW.
// =====================================================================================================================
constexpr int WG_KB = 64;        // voxels per pipeline stage
constexpr int WG_MAX_ASTAGES = 8;

struct UmmaWgradParams {
  CUtensorMap x_maps[8];
  // One kernel launch covers `nlaunch` independent sub-problems ("slots", blockIdx.z / nz): the tap groups of a
  // multi-group problem (each with its own dY lattice) -- or, in merge mode, runs of `merge_ng` groups side by side in N.
  CUtensorMap dy_maps[MTB200_MAX_GROUPS];         // dY lattice of group g (merge mode: slot l uses groups l*merge_ng ...)
  int merge_ng, merge_cout;                       // merge mode: groups side by side in N (N = ng * Cout), else 0
  int grp_widx[MTB200_MAX_GROUPS];                // merge mode: weight slice of group g
  int grp_off[MTB200_MAX_GROUPS][3];              // dY brick offset of group g
  int slot_tap_begin[MTB200_MAX_GROUPS];          // first tap of slot l (every slot has the same number of taps)
  int nz;                                         // N tiles per slot
  float* dw;
  int B, tiles_d, tiles_h, tiles_w, bd, bh, bw;
  long long nbricks;
  int bricks_per_cta;
  int Cin, Cout;            // padded channel counts
  int cbx, cby, BN;         // block widths (channels) of the X / dY boxes, N tile
  int blocks_per_tap;       // Cin / cbx
  int nblocks;              // (tap_end - tap_begin) * blocks_per_tap
  int blocks_per_tile;      // 128 / cbx
  int ntiles, tiles_per_cta;
  int astages, tmem_cols;
  int tap_map[MTB200_MAX_TAPS];
  int tap_coff[MTB200_MAX_TAPS][3];
  int tap_widx[MTB200_MAX_TAPS];
  int is_f16;
};

// MN-major descriptor: blocks of `cb` channels (swizzle span cb*2 bytes), LBO = distance between blocks along M/N,
// SBO = distance between 8-row (voxel) groups
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t cb_bytes, uint32_t lbo) {
  const uint32_t sbo = 8u * cb_bytes;
  const uint64_t layout = cb_bytes == 128 ? 2ull : (cb_bytes == 64 ? 4ull : 6ull);
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= layout << 61;
  return d;
}

__global__ void __launch_bounds__(UM_THREADS, 1) wgrad_taps_umma_kernel(const __grid_constant__ UmmaWgradParams p) {
  pdl_wait();  // programmatic dependent launch: nothing of the previous kernel is touched before this
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t a_full[WG_MAX_ASTAGES], a_empty[WG_MAX_ASTAGES];
  __shared__ __align__(8) uint64_t b_full[2], b_empty[2];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_stage_bytes = WG_KB * 128 * 2;                         // 16 KB: [64 voxels][128 rows]
  const uint32_t b_stage_bytes = ((WG_KB * p.BN * 2 + 1023) / 1024) * 1024;
  uint8_t* a_base = dsmem;
  uint8_t* b_base = dsmem + (size_t)p.astages * a_stage_bytes;
  const uint32_t xblock_bytes = WG_KB * p.cbx * 2, yblock_bytes = WG_KB * p.cby * 2;

  const long long brick0 = (long long)blockIdx.x * p.bricks_per_cta;
  const long long brick1 = min(p.nbricks, brick0 + p.bricks_per_cta);
  const int tile0 = blockIdx.y * p.tiles_per_cta;
  const int tile1 = min(p.ntiles, tile0 + p.tiles_per_cta);
  const int slot = (int)blockIdx.z / p.nz;
  const int n0 = ((int)blockIdx.z - slot * p.nz) * p.BN;
  const int tap_begin = p.slot_tap_begin[slot];
  const int nbricks = (int)max(0LL, brick1 - brick0);
  const int ntl = tile1 - tile0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.astages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (warp-uniform loop, one elected lane issues) =====
    if (nbricks > 0 && ntl > 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (long long br = brick0; br < brick1; ++br) {
        uint32_t t = (uint32_t)br;  // nbricks < 2^31: 32-bit divisions (the 64-bit chain was ~400 instructions per brick)
        const int tw = (int)(t % (uint32_t)p.tiles_w); t /= (uint32_t)p.tiles_w;
        const int th = (int)(t % (uint32_t)p.tiles_h); t /= (uint32_t)p.tiles_h;
        const int td = (int)(t % (uint32_t)p.tiles_d);
        const int b = (int)(t / (uint32_t)p.tiles_d);
        const int d0 = td * p.bd, h0 = th * p.bh, w0 = tw * p.bw;
        // dY brick (shared by every M-tile of this brick)
        mbar_wait(&b_empty[bs], bph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&b_full[bs], WG_KB * p.BN * 2);
          if (p.merge_ng) {  // one dY brick per tap group (its own output lattice), side by side along N
            const int nblk = p.merge_cout / p.cby;
            for (int j = 0; j < p.merge_ng; ++j) {
              const int g = slot * p.merge_ng + j;
              for (int h = 0; h < nblk; ++h)
                tma_load_5d(b_base + (size_t)bs * b_stage_bytes + (size_t)(j * nblk + h) * yblock_bytes, &p.dy_maps[g],
                            &b_full[bs], h * p.cby, w0 + p.grp_off[g][2], h0 + p.grp_off[g][1], d0 + p.grp_off[g][0], b);
            }
          } else {
            for (int j = 0; j < p.BN / p.cby; ++j)
              tma_load_5d(b_base + (size_t)bs * b_stage_bytes + (size_t)j * yblock_bytes, &p.dy_maps[slot], &b_full[bs],
                          n0 + j * p.cby, w0 + p.grp_off[slot][2], h0 + p.grp_off[slot][1], d0 + p.grp_off[slot][0], b);
          }
        }
        __syncwarp();
        if (++bs == 2) { bs = 0; bph ^= 1u; }
        // X row blocks, one stage per M-tile
        for (int tl = tile0; tl < tile1; ++tl) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&a_full[as], a_stage_bytes);
            for (int i = 0; i < p.blocks_per_tile; ++i) {
              int gb = tl * p.blocks_per_tile + i;
              if (gb >= p.nblocks) gb = 0;  // padding rows of the last tile: duplicate block 0 (never written back)
              const int tp = tap_begin + gb / p.blocks_per_tap;
              const int c0 = (gb % p.blocks_per_tap) * p.cbx;
              tma_load_5d(a_base + (size_t)as * a_stage_bytes + (size_t)i * xblock_bytes, &p.x_maps[p.tap_map[tp]],
                          &a_full[as], c0, w0 + p.tap_coff[tp][2], h0 + p.tap_coff[tp][1], d0 + p.tap_coff[tp][0], b);
            }
          }
          __syncwarp();
          if (++as == p.astages) { as = 0; aph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp-uniform loop, one elected lane issues =====
    if (nbricks > 0 && ntl > 0) {
      const uint32_t fmt = p.is_f16 ? 0u : 1u;
      // D=f32, A/B 16-bit, both MN-major (bits 15, 16), N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t cbx_bytes = p.cbx * 2, cby_bytes = p.cby * 2;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t a0 = __shfl_sync(0xffffffffu, (smem_u32(a_base) & 0x3FFFFu) >> 4, 0);
      const uint32_t b0 = __shfl_sync(0xffffffffu, (smem_u32(b_base) & 0x3FFFFu) >> 4, 0);
      // descriptor words precomputed: hi = SBO | version | swizzle, lo = (address >> 4) | LBO << 16
      const uint32_t lay_a = cbx_bytes == 128 ? 2u : (cbx_bytes == 64 ? 4u : 6u);
      const uint32_t lay_b = cby_bytes == 128 ? 2u : (cby_bytes == 64 ? 4u : 6u);
      const uint32_t hi_a = (((8u * cbx_bytes) >> 4) & 0x3FFFu) | (1u << 14) | (lay_a << 29);
      const uint32_t hi_b = (((8u * cby_bytes) >> 4) & 0x3FFFu) | (1u << 14) | (lay_b << 29);
      const uint32_t lbo_a = ((xblock_bytes >> 4) & 0x3FFFu) << 16, lbo_b = ((yblock_bytes >> 4) & 0x3FFFu) << 16;
      const uint32_t astage16 = a_stage_bytes >> 4, bstage16 = b_stage_bytes >> 4;
      const uint32_t ka = cbx_bytes, kb = cby_bytes;  // 16 voxels further along K = 16 * row bytes = (row bytes) x 16 B
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int br = 0; br < nbricks; ++br) {
        mbar_wait(&b_full[bs], bph);
        const uint32_t sb = (b0 + (uint32_t)bs * bstage16) | lbo_b;
        for (int tl = 0; tl < ntl; ++tl) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t sa = (a0 + (uint32_t)as * astage16) | lbo_a;
          if (elect_one()) {
            const uint32_t dcol = tmem_u + (uint32_t)(tl * p.BN);
#pragma unroll
            for (int k = 0; k < WG_KB / 16; ++k)
              umma_f16(dcol, ((uint64_t)hi_a << 32) | (uint64_t)(sa + k * ka), ((uint64_t)hi_b << 32) | (uint64_t)(sb + k * kb),
                       idesc, (br > 0 || k > 0) ? 1u : 0u);
            umma_commit(&a_empty[as]);
          }
          __syncwarp();
          if (++as == p.astages) { as = 0; aph ^= 1u; }
        }
        if (elect_one()) umma_commit(&b_empty[bs]);
        __syncwarp();
        if (++bs == 2) { bs = 0; bph ^= 1u; }
      }
      if (elect_one()) umma_commit(&acc_full);
      __syncwarp();
    }
  } else if (nbricks > 0 && ntl > 0) {
    // ===== epilogue: TMEM -> fp32 atomics into dW[widx][co][ci] =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    for (int tl = 0; tl < ntl; ++tl) {
      const int gb = (tile0 + tl) * p.blocks_per_tile + row / p.cbx;
      const bool valid = gb < p.nblocks;
      const int tp = tap_begin + (valid ? gb / p.blocks_per_tap : 0);
      const int ci = (valid ? (gb % p.blocks_per_tap) * p.cbx : 0) + row % p.cbx;
      float* dst = p.dw + ((long long)p.tap_widx[tp] * p.Cout + n0) * p.Cin + ci;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * p.BN + c0), r);
        if (valid) {
          float* dc = dst + (long long)c0 * p.Cin;
          if (p.merge_ng) {  // column block -> (tap group, channel)
            const int g = c0 / p.merge_cout;
            dc = p.dw + ((long long)p.grp_widx[slot * p.merge_ng + g] * p.Cout + (c0 - g * p.merge_cout)) * p.Cin + ci;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float v = __uint_as_float(r[j]);
            if (v != 0.f) atomicAdd(dc + (long long)j * p.Cin, v);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

static int block_width(int C) { return (C % 64 == 0) ? 64 : ((C % 32 == 0) ? 32 : 16); }

int wgrad_taps_umma(const mtb200_wgrad_params& p, cudaStream_t s) {
  if (!umma_available()) { set_error("wgrad_taps(umma): no sm_100 device"); return MTB200_ERR_UNSUPPORTED; }
  if (p.dtype != MTB200_BF16 && p.dtype != MTB200_F16) { set_error("wgrad_taps(umma): 16-bit only"); return MTB200_ERR_UNSUPPORTED; }
  if (p.xform) { set_error("wgrad_taps(umma): on-load transform not supported"); return MTB200_ERR_UNSUPPORTED; }
  if (p.Cin % 16 || p.Cout % 16 || p.in_ldc % 8 || p.in_coff % 8 || p.out_ldc % 8 || p.out_coff % 8) {
    set_error("wgrad_taps(umma): channel alignment"); return MTB200_ERR_UNSUPPORTED;
  }
  for (int k = 0; k < 3; ++k)
    if (p.is[k] < 1 || p.is[k] > 2 || p.os[k] < 1 || p.os[k] > 2) { set_error("wgrad_taps(umma): stride"); return MTB200_ERR_UNSUPPORTED; }
  const long long M = (long long)p.B * p.Do * p.Ho * p.Wo;
  if (M == 0) return MTB200_OK;
  if (p.impl != 3) {  // line-streaming kernels first (impl 3 = per-tap kernel only, impl 5 = line-streaming only)
    const int r2 = wgrad_line_s2_umma(p, s);  // stride-2 3x3x3 layers
    if (r2 != MTB200_ERR_UNSUPPORTED) return r2;
    const int r3 = wgrad_rows_umma(p, s);     // wide stride-1 layers below the second level (128-channel blocks)
    if (r3 != MTB200_ERR_UNSUPPORTED) return r3;
    const int r = wgrad_line_umma(p, s);
    if (r != MTB200_ERR_UNSUPPORTED) return r;
    if (p.impl == 5) { set_error("wgrad_taps(umma): problem outside the line-streaming kernel's envelope"); return r; }
  }
  EncodeTiledFn enc = get_encode_fn();
  const CUtensorMapDataType dt = p.dtype == MTB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;

  static thread_local UmmaWgradParams q;
  memset(&q, 0, sizeof(q));
  q.dw = p.dw;
  q.B = p.B; q.Cin = p.Cin; q.Cout = p.Cout;
  q.cbx = block_width(p.Cin);
  // Merge mode (ConvTranspose3d(k == s) weight gradient, generic_UNet.py:335-336): every tap group holds ONE tap with the
  // same input offset and differs only in its output lattice -> all groups share the X brick; their dY bricks sit side by
  // side along N (N = groups * Cout <= 256, `merge_per` groups per launch) and the activations are streamed once per
  // launch instead of once per group.
  int merge_per = (p.Cout == 16 || p.Cout == 32 || p.Cout == 64 || p.Cout == 128) ? min(p.ngroups, 256 / p.Cout) : 1;
  bool merge = p.ngroups > 1 && merge_per > 1 && p.ngroups % merge_per == 0;
  for (int g = 0; g < p.ngroups && merge; ++g) {
    merge = p.group_tap_begin[g + 1] - p.group_tap_begin[g] == 1;
    for (int k = 0; k < 3 && merge; ++k)
      merge = p.tap_off[p.group_tap_begin[g]][k] == p.tap_off[p.group_tap_begin[0]][k];
  }
  q.BN = merge ? merge_per * p.Cout : p.Cout;
  if (q.BN > 256) {
    q.BN = 0;
    for (int c = 256; c >= 16; c -= 16)
      if (p.Cout % c == 0) { q.BN = c; break; }
  }
  q.cby = merge ? min(p.Cout, 64) : block_width(q.BN);
  q.blocks_per_tap = p.Cin / q.cbx;
  q.blocks_per_tile = 128 / q.cbx;
  q.is_f16 = p.dtype == MTB200_F16;
  // voxel brick: 64 voxels, powers of two, minimise the number of bricks (ties: widest in w)
  long long best = -1;
  for (int bw = 64; bw >= 1; bw >>= 1)
    for (int bh = 64 / bw; bh >= 1; bh >>= 1) {
      const int bd = 64 / (bw * bh);
      const long long nt = (long long)((p.Do + bd - 1) / bd) * ((p.Ho + bh - 1) / bh) * ((p.Wo + bw - 1) / bw);
      if (best < 0 || nt < best) { best = nt; q.bd = bd; q.bh = bh; q.bw = bw; }
    }
  q.tiles_d = (p.Do + q.bd - 1) / q.bd; q.tiles_h = (p.Ho + q.bh - 1) / q.bh; q.tiles_w = (p.Wo + q.bw - 1) / q.bw;
  q.nbricks = (long long)p.B * q.tiles_d * q.tiles_h * q.tiles_w;

  // X parity maps (as in the forward kernel)
  int parity_slot[8];
  for (int i = 0; i < 8; ++i) parity_slot[i] = -1;
  int nmaps = 0;
  const int Dims[3] = {p.Di, p.Hi, p.Wi};
  const long long e = 2;
  for (int t = 0; t < p.ntaps; ++t) {
    int par[3], coord[3];
    for (int k = 0; k < 3; ++k) {
      const int off = p.tap_off[t][k];
      par[k] = ((off % p.is[k]) + p.is[k]) % p.is[k];
      coord[k] = floor_div(off - par[k], p.is[k]);
    }
    const int code = par[0] * 4 + par[1] * 2 + par[2];
    if (parity_slot[code] < 0) {
      cuuint64_t dims[5], strides[4];
      cuuint32_t box[5] = {(cuuint32_t)q.cbx, (cuuint32_t)q.bw, (cuuint32_t)q.bh, (cuuint32_t)q.bd, 1};
      long long ext[3];
      for (int k = 0; k < 3; ++k) {
        ext[k] = (Dims[k] - par[k] + p.is[k] - 1) / p.is[k];
        if (ext[k] < 1) { set_error("wgrad_taps(umma): empty parity lattice"); return MTB200_ERR_UNSUPPORTED; }
      }
      dims[0] = p.Cin; dims[1] = ext[2]; dims[2] = ext[1]; dims[3] = ext[0]; dims[4] = p.B;
      strides[0] = (cuuint64_t)p.in_ldc * e * p.is[2];
      strides[1] = (cuuint64_t)p.Wi * p.in_ldc * e * p.is[1];
      strides[2] = (cuuint64_t)p.Hi * p.Wi * p.in_ldc * e * p.is[0];
      strides[3] = (cuuint64_t)p.Di * p.Hi * p.Wi * p.in_ldc * e;
      uint8_t* base = (uint8_t*)p.x + ((((long long)par[0] * p.Hi + par[1]) * p.Wi + par[2]) * p.in_ldc + p.in_coff) * e;
      if (!encode_map(enc, &q.x_maps[nmaps], dt, 5, base, dims, strides, box, q.cbx * 2)) return MTB200_ERR_CUDA;
      parity_slot[code] = nmaps++;
    }
    q.tap_map[t] = parity_slot[code];
    for (int k = 0; k < 3; ++k) q.tap_coff[t][k] = coord[k];
    q.tap_widx[t] = p.tap_widx[t];
  }

  // pipeline / TMEM sizing
  const int max_tiles_tmem = 512 / q.BN;
  if (max_tiles_tmem < 1) { set_error("wgrad_taps(umma): BN too large"); return MTB200_ERR_UNSUPPORTED; }
  const int b_stage_bytes = ((WG_KB * q.BN * 2 + 1023) / 1024) * 1024;
  const int a_stage_bytes = WG_KB * 128 * 2;
  q.astages = max(2, min(WG_MAX_ASTAGES, (190 * 1024 - 2 * b_stage_bytes) / a_stage_bytes));
  const int smem = q.astages * a_stage_bytes + 2 * b_stage_bytes + 1024;
  cudaError_t ce = cudaFuncSetAttribute(wgrad_taps_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (ce != cudaSuccess) { set_error("wgrad_taps(umma): cudaFuncSetAttribute: %s", cudaGetErrorString(ce)); return MTB200_ERR_CUDA; }

  // dY lattice of every tap group: out coordinate = o * os + ooff
  {
    const int ODims[3] = {p.Dof, p.Hof, p.Wof};
    for (int g = 0; g < p.ngroups; ++g) {
      int par[3];
      cuuint64_t dims[5], strides[4];
      long long ext[3];
      for (int k = 0; k < 3; ++k) {
        const int off = p.group_ooff[g][k];
        par[k] = ((off % p.os[k]) + p.os[k]) % p.os[k];
        q.grp_off[g][k] = floor_div(off - par[k], p.os[k]);
        ext[k] = (ODims[k] - par[k] + p.os[k] - 1) / p.os[k];
        if (ext[k] < 1) { set_error("wgrad_taps(umma): empty dY lattice"); return MTB200_ERR_UNSUPPORTED; }
      }
      cuuint32_t box[5] = {(cuuint32_t)q.cby, (cuuint32_t)q.bw, (cuuint32_t)q.bh, (cuuint32_t)q.bd, 1};
      dims[0] = p.Cout; dims[1] = ext[2]; dims[2] = ext[1]; dims[3] = ext[0]; dims[4] = p.B;
      strides[0] = (cuuint64_t)p.out_ldc * e * p.os[2];
      strides[1] = (cuuint64_t)p.Wof * p.out_ldc * e * p.os[1];
      strides[2] = (cuuint64_t)p.Hof * p.Wof * p.out_ldc * e * p.os[0];
      strides[3] = (cuuint64_t)p.Dof * p.Hof * p.Wof * p.out_ldc * e;
      uint8_t* base = (uint8_t*)p.dy + ((((long long)par[0] * p.Hof + par[1]) * p.Wof + par[2]) * p.out_ldc + p.out_coff) * e;
      if (!encode_map(enc, &q.dy_maps[g], dt, 5, base, dims, strides, box, q.cby * 2)) return MTB200_ERR_CUDA;
      q.grp_widx[g] = p.group_tap_begin[g + 1] > p.group_tap_begin[g] ? p.tap_widx[p.group_tap_begin[g]] : 0;
    }
  }
  if (merge) { q.merge_ng = merge_per; q.merge_cout = p.Cout; }
  // Slots: one per tap group (merge mode: per run of `merge_per` groups).  Groups with the same number of taps share ONE
  // launch (grid.z = slots x N tiles) -- the transposed convolutions of the deep levels were 4 - 8 serialised launches
  // of ~30 CTAs each; a group without taps has nothing to do.  Groups with different tap counts: one launch each.
  const int per = merge ? merge_per : 1;
  const int nslots_all = p.ngroups / per;
  bool uniform = true;
  for (int g = 0; g < p.ngroups; ++g)
    uniform = uniform && (p.group_tap_begin[g + 1] - p.group_tap_begin[g]) == (p.group_tap_begin[1] - p.group_tap_begin[0]);
  // the kernel indexes dy_maps / grp_off by slot (or slot * merge_ng + j): a launch of a single slot l > 0 moves that
  // slot's entries to the front
  const UmmaWgradParams all = q;
  for (int l0 = 0; l0 < nslots_all; l0 += (uniform ? nslots_all : 1)) {
    const int nslots = uniform ? nslots_all : 1;
    for (int l = 0; l < nslots; ++l) {
      q.slot_tap_begin[l] = p.group_tap_begin[(l0 + l) * per];
      for (int j = 0; j < per; ++j) {
        const int gs = (l0 + l) * per + j, gd = l * per + j;
        q.dy_maps[gd] = all.dy_maps[gs];
        q.grp_widx[gd] = all.grp_widx[gs];
        for (int k = 0; k < 3; ++k) q.grp_off[gd][k] = all.grp_off[gs][k];
      }
    }
    const int ntaps_slot = p.group_tap_begin[l0 * per + 1] - p.group_tap_begin[l0 * per];
    if (ntaps_slot == 0) continue;
    q.nblocks = ntaps_slot * q.blocks_per_tap;
    q.ntiles = (q.nblocks + q.blocks_per_tile - 1) / q.blocks_per_tile;
    q.tiles_per_cta = min(q.ntiles, max_tiles_tmem);
    const int tile_groups = (q.ntiles + q.tiles_per_cta - 1) / q.tiles_per_cta;
    q.tmem_cols = 32;
    while (q.tmem_cols < q.tiles_per_cta * q.BN) q.tmem_cols *= 2;
    q.nz = merge ? 1 : p.Cout / q.BN;
    // split over voxel bricks so that the grid covers the machine about twice, with >= 4 bricks per CTA
    long long want = max(1LL, (2LL * num_sms()) / ((long long)tile_groups * q.nz * nslots));
    long long ksplit = min(want, max(1LL, q.nbricks / 4));
    q.bricks_per_cta = (int)((q.nbricks + ksplit - 1) / ksplit);
    ksplit = (q.nbricks + q.bricks_per_cta - 1) / q.bricks_per_cta;
    dim3 grid((unsigned)ksplit, tile_groups, q.nz * nslots);
    launch_pdl(wgrad_taps_umma_kernel, dim3(grid), dim3(UM_THREADS), (size_t)(smem), s, q);
    int r = check_launch("wgrad_taps_umma");
    if (r) return r;
  }
  return MTB200_OK;
}

EncodeTiledFn umma_encode_fn() { return get_encode_fn(); }

bool umma_encode_map(CUtensorMap* m, int dtype, int rank, void* base, const cuuint64_t* dims,
                     const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return false; }
  return encode_map(enc, m, dtype == MTB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                    rank, base, dims, strides_bytes, box, swizzle_bytes);
}

}  // namespace mtb
