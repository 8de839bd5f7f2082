// Sliding-window predictor: device-resident tile gather (flip-aware) and Gaussian-weighted scatter-add aggregation.
// Replaces the per-tile `.cpu().numpy()` + host numpy `+=` of neural_network.py:374-394 (739 MB D2H per tile in the
// reference) by fp32 accumulators that stay in HBM; algorithmic traffic = 8 B per (class, voxel) per tile (RMW).
#include "common.cuh"

namespace mtb {

__device__ __forceinline__ int flipc(int i, int n, int f) { return f ? n - 1 - i : i; }

// ---- tile gather ----------------------------------------------------------------------------------------------------
template <typename T>
__global__ void sw_gather_tile_kernel(const float* __restrict__ vol, int Cin, int X, int Y, int Z, int x0, int y0, int z0,
                                      int pd, int ph, int pw, int flip, T* __restrict__ tile, int ldc) {
  const int G = ldc / 8;
  const long long n = (long long)pd * ph * pw * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long v = i / G;
    const int w = (int)(v % pw); v /= pw;
    const int h = (int)(v % ph);
    const int d = (int)(v / ph);
    const int sx = x0 + flipc(d, pd, flip & 4), sy = y0 + flipc(h, ph, flip & 2), sz = z0 + flipc(w, pw, flip & 1);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      o[j] = c < Cin ? vol[(((long long)c * X + sx) * Y + sy) * Z + sz] : 0.f;
    }
    store8<T>(tile + (i / G) * ldc + cg * 8, o);
  }
}

int sw_gather_tile(const float* vol, int Cin, int X, int Y, int Z, int x0, int y0, int z0, int pd, int ph, int pw,
                   int flip, void* tile, int dtype, int ldc, cudaStream_t s) {
  MTB_REQUIRE(ldc % 8 == 0 && Cin <= ldc, "sw_gather_tile: ldc=%d Cin=%d", ldc, Cin);
  MTB_REQUIRE(x0 >= 0 && y0 >= 0 && z0 >= 0 && x0 + pd <= X && y0 + ph <= Y && z0 + pw <= Z,
              "sw_gather_tile: tile outside the volume");
  const long long n = (long long)pd * ph * pw * (ldc / 8);
  const int blocks = (int)min((long long)num_sms() * 8, (n + 255) / 256);
  if (blocks == 0) return MTB200_OK;
  MTB_DISPATCH_DTYPE(dtype, T, (sw_gather_tile_kernel<T><<<blocks, 256, 0, s>>>(vol, Cin, X, Y, Z, x0, y0, z0, pd, ph, pw,
                                                                                flip, reinterpret_cast<T*>(tile), ldc)));
  return check_launch("sw_gather_tile");
}

// ---- aggregation -----------------------------------------------------------------------------------------------------
// One block = one (d, h) row segment of SEG consecutive w voxels x all channels.  Phase 1 reads the channels-last logits
// (coalesced), applies the inference non-linearity * weight * gauss, and parks them transposed in smem; phase 2 does the
// read-modify-write on the channels-first accumulator with 128-byte coalesced rows.  Tiles of one launch never overlap
// => plain RMW.  nonlin: 0 = none, 1 = sigmoid (MultiTalent, MultiTalent_Trainer_DDP.py:46), 2 = softmax over the C
// channels (`softmax_helper` of the single-task trainers, nnUNetTrainerV2.py:162): the raw logits are parked first and
// one thread per voxel normalises its column of the shared tile.
constexpr int SEG = 128;  // 64 measured 3.2 TB/s on the RMW (256-byte rows per class plane); 128 = whole 512-byte patch rows
constexpr int AGG_T = 256;

template <typename T>
__global__ void __launch_bounds__(AGG_T) sw_aggregate_kernel(const T* __restrict__ logits, int ldc, int C, int pd, int ph,
                                                             int pw, int flip, const float* __restrict__ gauss,
                                                             float weight, int apply_sigmoid, float* __restrict__ acc,
                                                             float* __restrict__ nb, int X, int Y, int Z, int x0, int y0,
                                                             int z0) {
  extern __shared__ float sm[];  // [Cp8][SEG+1]
  const int nseg = (pw + SEG - 1) / SEG;
  const int seg = blockIdx.x % nseg;
  const int h = (blockIdx.x / nseg) % ph;
  const int d = blockIdx.x / (nseg * ph);
  const int w0 = seg * SEG;
  const int nw = min(SEG, pw - w0);
  const int G = (C + 7) / 8;
  const int sd = flipc(d, pd, flip & 4), sh = flipc(h, ph, flip & 2);
  // phase 1
  for (int i = threadIdx.x; i < nw * G; i += AGG_T) {
    const int cg = i % G, wl = i / G;
    const int w = w0 + wl;
    const int sw = flipc(w, pw, flip & 1);
    float z[8];
    load8<T>(logits + (((long long)sd * ph + sh) * pw + sw) * ldc + cg * 8, z);
    const float gw = apply_sigmoid == 2 ? 1.f : weight * (gauss ? gauss[((long long)d * ph + h) * pw + w] : 1.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = z[j];
      if (apply_sigmoid == 1) {
        const float e = expf(-fabsf(v));
        v = v >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
      }
      sm[(cg * 8 + j) * (SEG + 1) + wl] = v * gw;
    }
  }
  __syncthreads();
  if (apply_sigmoid == 2) {  // softmax over the channels of each voxel (column wl of the shared tile)
    for (int wl = threadIdx.x; wl < nw; wl += AGG_T) {
      float m = -INFINITY;
      for (int c = 0; c < C; ++c) m = fmaxf(m, sm[c * (SEG + 1) + wl]);
      float ssum = 0.f;
      for (int c = 0; c < C; ++c) {
        const float e = expf(sm[c * (SEG + 1) + wl] - m);
        sm[c * (SEG + 1) + wl] = e;
        ssum += e;
      }
      const float gw = weight * (gauss ? gauss[((long long)d * ph + h) * pw + w0 + wl] : 1.f) / ssum;
      for (int c = 0; c < C; ++c) sm[c * (SEG + 1) + wl] *= gw;
    }
    __syncthreads();
  }
  // phase 2
  const long long plane = (long long)Y * Z;
  const long long rowbase = ((long long)(x0 + d) * Y + (y0 + h)) * Z + z0 + w0;
  // all loads of a batch are issued before the first store: the compiler cannot reorder a load above an earlier store
  // through a different pointer, so a plain `*a += v` loop serialises one DRAM round trip per element
  constexpr int RB = 8;
  for (int i0 = threadIdx.x; i0 < C * SEG; i0 += AGG_T * RB) {
    float old[RB];
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const int i = i0 + k * AGG_T;
      const int c = i / SEG, wl = i % SEG;
      old[k] = (i < C * SEG && wl < nw) ? acc[(long long)c * X * plane + rowbase + wl] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < RB; ++k) {
      const int i = i0 + k * AGG_T;
      const int c = i / SEG, wl = i % SEG;
      if (i < C * SEG && wl < nw) acc[(long long)c * X * plane + rowbase + wl] = old[k] + sm[c * (SEG + 1) + wl];
    }
  }
  if (nb) {
    for (int wl = threadIdx.x; wl < nw; wl += AGG_T)
      nb[rowbase + wl] += gauss ? gauss[((long long)d * ph + h) * pw + w0 + wl] : 1.f;
  }
}

int sw_aggregate(const void* logits, int dtype, int ldc, int C, int pd, int ph, int pw, int flip, const float* gauss,
                 float weight, int apply_sigmoid, float* acc, float* nb, int X, int Y, int Z, int x0, int y0, int z0,
                 cudaStream_t s) {
  MTB_REQUIRE(ldc % 8 == 0 && ((C + 7) / 8) * 8 <= ldc, "sw_aggregate: C=%d ldc=%d", C, ldc);
  MTB_REQUIRE(x0 >= 0 && y0 >= 0 && z0 >= 0 && x0 + pd <= X && y0 + ph <= Y && z0 + pw <= Z,
              "sw_aggregate: tile outside the volume");
  const int nseg = (pw + SEG - 1) / SEG;
  const long long blocks = (long long)pd * ph * nseg;
  if (blocks == 0) return MTB200_OK;
  const size_t smem = (size_t)((C + 7) / 8) * 8 * (SEG + 1) * sizeof(float);
  MTB_DISPATCH_DTYPE(dtype, T, (sw_aggregate_kernel<T><<<(unsigned)blocks, AGG_T, smem, s>>>(
      reinterpret_cast<const T*>(logits), ldc, C, pd, ph, pw, flip, gauss, weight, apply_sigmoid, acc, nb, X, Y, Z, x0,
      y0, z0)));
  return check_launch("sw_aggregate");
}

// ---- finalize: normalise + threshold ----------------------------------------------------------------------------------
__global__ void sw_finalize_kernel(float* __restrict__ acc, const float* __restrict__ nb, int C, long long nvox,
                                   const float* __restrict__ class_order, float* __restrict__ seg, long long cstride) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
    const float n = nb[v];
    float sv = 0.f, best = -INFINITY;
    for (int c = 0; c < C; ++c) {
      const float p = acc[(long long)c * cstride + v] / n;  // IEEE division, as numpy's in-place /=
      acc[(long long)c * cstride + v] = p;
      if (class_order) {
        if (p > 0.5f) sv = class_order[c];
      } else if (p > best) {
        best = p;
        sv = (float)c;
      }
    }
    if (seg) seg[v] = sv;
  }
}

int sw_finalize(float* acc, const float* nb, int C, long long cstride, long long nvox, const float* class_order, float* seg,
                cudaStream_t s) {
  const int blocks = (int)min((long long)num_sms() * 8, (nvox + 255) / 256);
  if (blocks == 0) return MTB200_OK;
  sw_finalize_kernel<<<blocks, 256, 0, s>>>(acc, nb, C, nvox, class_order, seg, cstride);
  return check_launch("sw_finalize");
}

}  // namespace mtb
