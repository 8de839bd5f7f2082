// Shared device/host helpers for libmtb200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mtb200.h"

namespace mtb {

// ---- error plumbing: never abort/throw across the C boundary ----------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> status

#define MTB_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      mtb::set_error(__VA_ARGS__);      \
      return MTB200_ERR_INVALID;        \
    }                                   \
  } while (0)

// ---- storage-type traits ------------------------------------------------------------------------------------------
template <typename T> struct Traits;
template <> struct Traits<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  static __device__ __forceinline__ float round(float v) { return v; }
};
template <> struct Traits<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
};
template <> struct Traits<__half> {
  static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
  static __device__ __forceinline__ float round(float v) { return __half2float(__float2half_rn(v)); }
};

// load / store 8 consecutive elements (16-byte aligned for 16-bit types, 32-byte for float) as floats
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}
template <> __device__ __forceinline__ void store8<__half>(__half* p, const float (&v)[8]) {
  uint4 r;
  __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}

// 8 consecutive elements kept in their storage format (4 registers for the 16-bit types): lets a streaming kernel
// keep several vectors in flight without paying 8 fp32 registers per vector
template <typename T> struct Raw8;
// `load` is an asm volatile read-only-path load: the compiler keeps volatile asm statements in program order, so a batch
// of loads written before the arithmetic is really issued before it (plain loads get sunk next to their first use, which
// serialises one DRAM round trip per vector -- measured on in_bwd_apply: 3.8 -> 5.6 TB/s)
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4));
  }
  // coherent variant (no .nc; streaming, not kept in L1): for an operand the same kernel also writes (in-place passes)
  __device__ __forceinline__ void loadc(const float* p) {
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4));
  }
  __device__ __forceinline__ float get(int i) const {
    return i == 0 ? a.x : i == 1 ? a.y : i == 2 ? a.z : i == 3 ? a.w : i == 4 ? b.x : i == 5 ? b.y : i == 6 ? b.z : b.w;
  }
};
template <> struct Raw8<__nv_bfloat16> {
  uint4 r;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  }
  __device__ __forceinline__ void loadc(const __nv_bfloat16* p) {
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  }
  __device__ __forceinline__ float get(int i) const {
    const uint32_t w = i < 2 ? r.x : i < 4 ? r.y : i < 6 ? r.z : r.w;
    return __uint_as_float((i & 1) ? (w & 0xFFFF0000u) : (w << 16));
  }
};
template <> struct Raw8<__half> {
  uint4 r;
  __device__ __forceinline__ void load(const __half* p) {
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  }
  __device__ __forceinline__ void loadc(const __half* p) {
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  }
  __device__ __forceinline__ float get(int i) const {
    const uint32_t w = i < 2 ? r.x : i < 4 ? r.y : i < 6 ? r.z : r.w;
    const __half2 h = *reinterpret_cast<const __half2*>(&w);
    return (i & 1) ? __high2float(h) : __low2float(h);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dispatch on the activation dtype enum
#define MTB_DISPATCH_DTYPE(dt, T, ...)                                        \
  switch (dt) {                                                               \
    case MTB200_F32: { using T = float; __VA_ARGS__; break; }                 \
    case MTB200_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }        \
    case MTB200_F16: { using T = __half; __VA_ARGS__; break; }                \
    default: mtb::set_error("bad dtype %d", (int)(dt)); return MTB200_ERR_INVALID; \
  }

int num_sms();

// ---- programmatic dependent launch -----------------------------------------------------------------------------------
// A training step is ~190 dependent launches in one stream; measured on the B200 (tools/pdl_probe.cu) a dependent launch
// costs 3.9 us of idle time between two kernels, 1.6 us when the second kernel is launched with the programmatic stream
// serialization attribute (its launch is processed while the first one still runs) and begins with griddepcontrol.wait
// (= wait until every earlier kernel of the stream has finished and its writes are visible).  `pdl_wait()` is the FIRST
// statement of every kernel launched through `launch_pdl`; MTB200_PDL=0 launches without the attribute (the wait is then a
// no-op).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();  // api.cu
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// the same with thread-block clusters of `cluster_x` consecutive CTAs (grid.x must be a multiple)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                      unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace mtb
