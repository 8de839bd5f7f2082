// Plane-streaming ("halo-resident") tcgen05 convolution for stride-1 kernels with taps in [-1,1]^3 and Cin_p <= 64.
//
// The per-tap kernel (conv_umma.cu) re-fetches every activation tile once per tap: 27x L2->SM traffic, which saturates
// the L2 path at full resolution (N = 32) long before the tensor pipe.  Here a CTA owns an output column
//     16 h-lines x (8*TW) w-columns x a segment of d-planes
// and streams the INPUT planes through a ring in shared memory exactly once: plane p = one 5-D TMA box
// [18 h][8*TW+2 w][Cin_p] (halo included, out-of-volume = zero fill = conv padding).  A tap (dz,dy,dx) of output plane d
// is then just a shifted VIEW of ring slot (d+dz): the K-major UMMA descriptor starts at row (dy+1)*Wh + (dx+1) + 8*tw
// with an 8-row-group stride of Wh rows.  (tools/umma_probe.cu showed on the B200 that descriptors with any row-granular
// start / group stride work with base_offset = 0.)  L2->SM traffic drops from 27x to (18/16)*(Wh/(8*TW)) ~ 1.2-1.4x.
//
// Warp roles (7 warps): 0 = plane producer (TMA), 1 = TMEM owner + MMA issuer, 2 = weight-tile producer (TMA),
// 3..6 = epilogue.  TMEM holds 2 x TW accumulators of BN columns: the epilogue of plane d overlaps the MMAs of d+1.
#include "umma.cuh"

namespace mtb {

using namespace um;

constexpr int HL_THREADS = 224;
constexpr int HL_RING = 4;        // input planes resident
constexpr int HL_MAX_WSTAGES = 8;
constexpr int HL_HT = 16;         // output h-lines per CTA
constexpr int HL_HH = HL_HT + 2;

struct HaloParams {
  CUtensorMap a_map, w_map;
  void* out;
  const float* bias;
  double* stats;
  int B, D, H, W;             // output grid == input grid (stride 1)
  int out_ldc, out_coff, Cout;
  int TW, Wh, BN, pitch;      // w sub-tiles of 8, halo width, N tile, bytes per row (= Cin_p * 2)
  int plane_bytes, plane_tx, wtile_bytes, wstages, tmem_cols;
  int tiles_h, tiles_w, nseg, seglen;
  int ntaps;
  int tap_off[MTB200_MAX_TAPS][3];
  int tap_widx[MTB200_MAX_TAPS];
  int accumulate, is_f16;
};

template <typename T>
__global__ void __launch_bounds__(HL_THREADS, 1) conv_halo_umma_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t dsmem_raw[];
  __shared__ __align__(8) uint64_t plane_full[HL_RING], plane_empty[HL_RING];
  __shared__ __align__(8) uint64_t w_full[HL_MAX_WSTAGES], w_empty[HL_MAX_WSTAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_sum[128], s_sq[128];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* dsmem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* plane_base = dsmem;
  uint8_t* w_base = dsmem + (size_t)HL_RING * p.plane_bytes;

  // work unit
  int u = blockIdx.x;
  const int seg = u % p.nseg; u /= p.nseg;
  const int twg = u % p.tiles_w; u /= p.tiles_w;
  const int th = u % p.tiles_h;
  const int b = u / p.tiles_h;
  const int h0 = th * HL_HT, w0 = twg * 8 * p.TW;
  const int ds = seg * p.seglen, de = min(p.D, ds + p.seglen);
  const int nout = de - ds;            // output planes of this CTA (>= 1 by construction)
  const int n0 = blockIdx.y * p.BN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < HL_RING; ++i) { mbar_init(&plane_full[i], 1); mbar_init(&plane_empty[i], 1); }
    for (int i = 0; i < p.wstages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 128) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; }
  if (warp == 1) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== input-plane producer: planes ds-1 .. de =====
    if (lane == 0) {
      const int nplanes = nout + 2;
      for (int j = 0; j < nplanes; ++j) {
        const int slot = j % HL_RING;
        mbar_wait(&plane_empty[slot], (((uint32_t)(j / HL_RING)) & 1u) ^ 1u);
        mbar_expect_tx(&plane_full[slot], (uint32_t)p.plane_tx);
        tma_load_5d(plane_base + (size_t)slot * p.plane_bytes, &p.a_map, &plane_full[slot], 0, w0 - 1, h0 - 1,
                    ds - 1 + j, b);
      }
    }
  } else if (warp == 2) {
    // ===== weight-tile producer: ntaps tiles per output plane =====
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0;
      for (int i = 0; i < nout; ++i)
        for (int t = 0; t < p.ntaps; ++t) {
          mbar_wait(&w_empty[ws], wph ^ 1u);
          mbar_expect_tx(&w_full[ws], (uint32_t)(p.BN * p.pitch));
          tma_load_3d(w_base + (size_t)ws * p.wtile_bytes, &p.w_map, &w_full[ws], 0, n0, p.tap_widx[t]);
          if (++ws == p.wstages) { ws = 0; wph ^= 1u; }
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = idesc_f16(p.is_f16 != 0, (uint32_t)p.BN, false, false);
      const uint32_t sbo_a = (uint32_t)(p.Wh * p.pitch), sbo_b = 8u * p.pitch;
      const int ksteps = p.pitch / 32;
      int ws = 0;
      uint32_t wph = 0;
      for (int i = 0; i < nout; ++i) {
        const int buf = i & 1;
        mbar_wait(&acc_empty[buf], (((uint32_t)(i >> 1)) & 1u) ^ 1u);
        // planes needed: j = i (d-1), i+1 (d), i+2 (d+1); only the newest has not been waited for yet
        for (int j = (i == 0 ? 0 : i + 2); j <= i + 2; ++j)
          mbar_wait(&plane_full[j % HL_RING], ((uint32_t)(j / HL_RING)) & 1u);
        tc_fence_after();
        for (int t = 0; t < p.ntaps; ++t) {
          mbar_wait(&w_full[ws], wph);
          tc_fence_after();
          const int slot = (i + 1 + p.tap_off[t][0]) % HL_RING;
          const uint32_t a0 = smem_u32(plane_base + (size_t)slot * p.plane_bytes) +
                              (uint32_t)(((p.tap_off[t][1] + 1) * p.Wh + (p.tap_off[t][2] + 1)) * p.pitch);
          const uint32_t b0 = smem_u32(w_base + (size_t)ws * p.wtile_bytes);
          for (int tw = 0; tw < p.TW; ++tw) {
            const uint32_t dcol = tmem_base + (uint32_t)((buf * p.TW + tw) * p.BN);
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t da = kmajor_desc(a0 + (uint32_t)(tw * 8 * p.pitch + k * 32), (uint32_t)p.pitch, sbo_a);
              const uint64_t db = kmajor_desc(b0 + (uint32_t)(k * 32), (uint32_t)p.pitch, sbo_b);
              umma_f16(dcol, da, db, idesc, (t > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&w_empty[ws]);
          if (++ws == p.wstages) { ws = 0; wph ^= 1u; }
        }
        umma_commit(&acc_full[buf]);
        umma_commit(&plane_empty[i % HL_RING]);  // plane d-1 is not needed by later output planes
      }
    }
  } else {
    // ===== epilogue warps 3..6 (TMEM lane quarter = warp % 4) =====
    const int q = warp & 3;
    const int m = q * 32 + lane;          // row of the 128-row tile
    const int hh = h0 + (m >> 3);
    T* out = reinterpret_cast<T*>(p.out);
    float csum[8], csq[8];                // this lane's column partials, one per 16-column chunk
#pragma unroll
    for (int c = 0; c < 8; ++c) { csum[c] = 0.f; csq[c] = 0.f; }
    for (int i = 0; i < nout; ++i) {
      const int buf = i & 1;
      const int d = ds + i;
      mbar_wait(&acc_full[buf], ((uint32_t)(i >> 1)) & 1u);
      tc_fence_after();
      for (int tw = 0; tw < p.TW; ++tw) {
        const int ww = w0 + tw * 8 + (m & 7);
        const bool valid = hh < p.H && ww < p.W;
        T* orow = out + ((((long long)b * p.D + d) * p.H + hh) * p.W + ww) * p.out_ldc + p.out_coff + n0;
        const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * p.TW + tw) * p.BN);
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const int c0 = cc * 16;
          if (c0 < p.BN) {
            uint32_t r[16];
            float v[16];
            tmem_ld16(tcol + (uint32_t)c0, r);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += __ldg(p.bias + n0 + c0 + j);
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
              float lo[8], hi[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
              store8<T>(orow + c0, lo);
              store8<T>(orow + c0 + 8, hi);
            }
            if (p.stats) {
              float s[16], ss[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float x = valid ? Traits<T>::round(v[j]) : 0.f;
                s[j] = x;
                ss[j] = x * x;
              }
              warp_colsum16(s, lane);
              warp_colsum16(ss, lane);
              csum[cc] += s[0];
              csq[cc] += ss[0];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if (p.stats && (lane & 1) == 0) {
      const int col = colsum16_column(lane);
#pragma unroll
      for (int cc = 0; cc < 8; ++cc)
        if (cc * 16 < p.BN) {
          atomicAdd(&s_sum[cc * 16 + col], csum[cc]);
          atomicAdd(&s_sq[cc * 16 + col], csq[cc]);
        }
    }
  }
  __syncthreads();
  if (p.stats) {
    for (int c = threadIdx.x; c < p.BN; c += HL_THREADS) {
      if (s_sum[c] != 0.f || s_sq[c] != 0.f) {
        double* st = p.stats + ((long long)b * p.Cout + n0 + c) * 2;
        atomicAdd(st, (double)s_sum[c]);
        atomicAdd(st + 1, (double)s_sq[c]);
      }
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// Returns MTB200_ERR_UNSUPPORTED when the problem is outside this kernel's envelope (caller falls back to the per-tap
// kernel); all generic argument checks were done by the caller.
int conv_halo_umma(const mtb200_conv_params& p, cudaStream_t s) {
  if (p.ngroups != 1) return MTB200_ERR_UNSUPPORTED;
  for (int k = 0; k < 3; ++k)
    if (p.is[k] != 1 || p.os[k] != 1 || p.group_ooff[0][k] != 0) return MTB200_ERR_UNSUPPORTED;
  if (p.Do != p.Di || p.Ho != p.Hi || p.Wo != p.Wi || p.Dof != p.Do || p.Hof != p.Ho || p.Wof != p.Wo)
    return MTB200_ERR_UNSUPPORTED;
  if (p.Cin != 16 && p.Cin != 32 && p.Cin != 64) return MTB200_ERR_UNSUPPORTED;
  bool spatial = false;
  for (int t = 0; t < p.ntaps; ++t)
    for (int k = 0; k < 3; ++k) {
      if (p.tap_off[t][k] < -1 || p.tap_off[t][k] > 1) return MTB200_ERR_UNSUPPORTED;
      if (p.tap_off[t][k] != 0) spatial = true;
    }
  if (!spatial || p.ntaps < 9) return MTB200_ERR_UNSUPPORTED;  // 1x1x1: nothing to reuse
  if (p.Ho < 8 || p.Wo < 8) return MTB200_ERR_UNSUPPORTED;      // tiny maps: tiles would be mostly padding

  static HaloParams q;
  memset(&q, 0, sizeof(q));
  q.pitch = p.Cin * 2;
  q.BN = p.Cout;
  if (q.BN > 128) {
    q.BN = 0;
    for (int c = 128; c >= 16; c -= 16)
      if (p.Cout % c == 0) { q.BN = c; break; }
  }
  if (q.BN == 0) return MTB200_ERR_UNSUPPORTED;
  q.wtile_bytes = ((q.BN * q.pitch + 1023) / 1024) * 1024;
  const int smem_budget = 200 * 1024;
  q.TW = 0;
  for (int tw = 4; tw >= 1; tw >>= 1) {
    if (tw > 1 && 8 * (tw / 2) >= p.Wo) continue;  // do not tile wider than the map needs
    const int wh = 8 * tw + 2;
    const int pb = ((HL_HH * wh * q.pitch + 1023) / 1024) * 1024;
    if (2 * tw * q.BN > 512) continue;
    if (HL_RING * pb + 2 * q.wtile_bytes > smem_budget) continue;
    q.TW = tw; q.Wh = wh; q.plane_bytes = pb;
    break;
  }
  if (q.TW == 0) return MTB200_ERR_UNSUPPORTED;
  q.plane_tx = HL_HH * q.Wh * q.pitch;
  q.wstages = max(2, min(HL_MAX_WSTAGES, (smem_budget - HL_RING * q.plane_bytes) / q.wtile_bytes));
  q.tmem_cols = 32;
  while (q.tmem_cols < 2 * q.TW * q.BN) q.tmem_cols *= 2;

  // tensor maps: activations [C][W][H][D][B] (box = one halo plane), weights [Cin][Cout][taps]
  {
    cuuint64_t dims[5] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Wi, (cuuint64_t)p.Hi, (cuuint64_t)p.Di, (cuuint64_t)p.B};
    cuuint64_t strides[4] = {(cuuint64_t)p.in_ldc * 2, (cuuint64_t)p.Wi * p.in_ldc * 2,
                             (cuuint64_t)p.Hi * p.Wi * p.in_ldc * 2, (cuuint64_t)p.Di * p.Hi * p.Wi * p.in_ldc * 2};
    cuuint32_t box[5] = {(cuuint32_t)p.Cin, (cuuint32_t)q.Wh, (cuuint32_t)HL_HH, 1, 1};
    if (!umma_encode_map(&q.a_map, p.dtype, 5, (uint8_t*)p.in + (size_t)p.in_coff * 2, dims, strides, box, q.pitch))
      return MTB200_ERR_CUDA;
  }
  {
    int n_widx = 0;
    for (int t = 0; t < p.ntaps; ++t) n_widx = max(n_widx, p.tap_widx[t] + 1);
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, (cuuint64_t)n_widx};
    cuuint64_t strides[2] = {(cuuint64_t)p.Cin * 2, (cuuint64_t)p.Cin * p.Cout * 2};
    cuuint32_t box[3] = {(cuuint32_t)p.Cin, (cuuint32_t)q.BN, 1};
    if (!umma_encode_map(&q.w_map, p.dtype, 3, (void*)p.w, dims, strides, box, q.pitch)) return MTB200_ERR_CUDA;
  }
  q.out = p.out; q.bias = p.bias; q.stats = p.stats;
  q.B = p.B; q.D = p.Do; q.H = p.Ho; q.W = p.Wo;
  q.out_ldc = p.out_ldc; q.out_coff = p.out_coff; q.Cout = p.Cout;
  q.ntaps = p.ntaps;
  for (int t = 0; t < p.ntaps; ++t) {
    for (int k = 0; k < 3; ++k) q.tap_off[t][k] = p.tap_off[t][k];
    q.tap_widx[t] = p.tap_widx[t];
  }
  q.accumulate = p.accumulate;
  q.is_f16 = p.dtype == MTB200_F16;
  q.tiles_h = (p.Ho + HL_HT - 1) / HL_HT;
  q.tiles_w = (p.Wo + 8 * q.TW - 1) / (8 * q.TW);
  const int ny = p.Cout / q.BN;
  const long long cols = (long long)p.B * q.tiles_h * q.tiles_w * ny;
  // split D into segments so that the grid covers the machine ~4x (each segment re-loads 2 halo planes)
  long long want = (4LL * num_sms() + cols - 1) / cols;
  q.nseg = (int)max(1LL, min(want, (long long)max(1, p.Do / 8)));
  q.seglen = (p.Do + q.nseg - 1) / q.nseg;
  q.nseg = (p.Do + q.seglen - 1) / q.seglen;
  const long long units = (long long)p.B * q.tiles_h * q.tiles_w * q.nseg;
  MTB_REQUIRE(units < (1LL << 31), "conv_halo: too many work units");

  const int smem = HL_RING * q.plane_bytes + q.wstages * q.wtile_bytes + 1024;
  dim3 grid((unsigned)units, ny, 1);
  cudaError_t e;
  if (p.dtype == MTB200_BF16) {
    e = cudaFuncSetAttribute(conv_halo_umma_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) conv_halo_umma_kernel<__nv_bfloat16><<<grid, HL_THREADS, smem, s>>>(q);
  } else {
    e = cudaFuncSetAttribute(conv_halo_umma_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) conv_halo_umma_kernel<__half><<<grid, HL_THREADS, smem, s>>>(q);
  }
  if (e != cudaSuccess) { set_error("conv_halo: cudaFuncSetAttribute(%d B): %s", smem, cudaGetErrorString(e)); return MTB200_ERR_CUDA; }
  return check_launch("conv_halo_umma");
}

}  // namespace mtb
