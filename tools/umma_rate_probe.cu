// Probe: issue-to-completion cost of one tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16, both operands in
// shared memory) as a function of N, of the operand layout (K-major SW128 / SW64 / SW32, MN-major SW128) and of whether
// consecutive MMAs read DIFFERENT A tiles (what an implicit-GEMM conv does: every tap is another shifted view) or the
// same one.  One CTA per SM, one issuing lane, 64 MMAs per commit; reports cycles per MMA (clock64 on the SM) and the
// chip-level TFLOP/s that rate corresponds to.  Answers "what does a wider N buy" for the conv kernels' tile choices.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_rate_probe tools/umma_rate_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

struct Rate {
  int n;            // MMA N
  int row_bytes;    // K-major: swizzle span = row pitch (128 / 64 / 32); MN-major: 128
  int mn_major;     // 1: both operands MN-major (wgrad-style), SW128
  int distinct_a;   // 1: A start address moves by one row per MMA and by 16 KB blocks (conv taps); 0: fixed
  int distinct_b;   // 1: B start moves too
  int sbo_rows;     // K-major A operand: rows between consecutive 8-row groups (8 = dense; 8*TW+2 = a halo tile's line pitch)
  int rounds;       // commits
  long long* clk;   // [gridDim.x]
};

__global__ void __launch_bounds__(128, 1) rate_kernel(const Rate p) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t done_bar[2];
  __shared__ uint32_t tmem_slot;
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  // fill with small finite bf16 values (realistic switching activity, no NaN/denormal side effects)
  for (int i = threadIdx.x; i < 200 * 1024 / 2; i += blockDim.x)
    reinterpret_cast<__nv_bfloat16*>(sm)[i] = __float2bfloat16(((i * 37 + 11) % 29 - 14) * 0.03125f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&done_bar[0], 1); mbar_init(&done_bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  if (warp == 0) {
    // issue loop shaped like the conv kernels': warp-uniform control flow, one elected lane issues, descriptor low
    // words precomputed in (uniform) registers, 16 MMAs fully unrolled per commit group
    const uint32_t a0 = __shfl_sync(0xffffffffu, (smem_u32(sm) & 0x3FFFFu) >> 4, 0);
    const uint32_t b0 = a0 + ((128u * 1024u) >> 4);
    const uint32_t layout = p.row_bytes == 128 ? 2u : (p.row_bytes == 64 ? 4u : 6u);
    uint32_t hi_a, hi_b, lbo_a = 0, lbo_b = 0;
    if (p.mn_major) {  // 64-byte rows = 32 channels, rows = K; A blocks = row-shifted views, B blocks = 8 KB lines
      hi_a = hi_b = ((8u * 64u) >> 4) | (1u << 14) | (4u << 29);
      lbo_a = (64u >> 4) << 16;
      lbo_b = (8192u >> 4) << 16;
    } else {
      hi_b = (((uint32_t)(8 * p.row_bytes)) >> 4) | (1u << 14) | (layout << 29);
      hi_a = (((uint32_t)(p.sbo_rows * p.row_bytes)) >> 4) | (1u << 14) | (layout << 29);
    }
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((p.mn_major ? 1u : 0u) << 15) | ((p.mn_major ? 1u : 0u) << 16) |
                           (((uint32_t)p.n >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t ksteps = p.mn_major ? 8u : (uint32_t)p.row_bytes / 32u;
    uint32_t a_off[16], b_off[16];
#pragma unroll
    for (uint32_t g = 0; g < 16; ++g) {
      const uint32_t k = g % ksteps, t = g / ksteps;
      uint32_t a = p.mn_major ? k * 1024u : k * 32u, b = a;
      if (p.distinct_a) a += (t % 3) * (uint32_t)p.row_bytes + ((t / 3) % 3) * 32768u;
      if (p.distinct_b && !p.mn_major) b += (t % 9) * 4096u;
      a_off[g] = (a0 + (a >> 4)) | lbo_a;
      b_off[g] = (b0 + (b >> 4)) | lbo_b;
    }
    const long long t0 = clock64();
    for (int r = 0; r < p.rounds; ++r) {
      if (r >= 2) mbar_wait(&done_bar[r & 1], (uint32_t)(((r - 2) >> 1) & 1));  // two commit groups in flight
      if (elect_one()) {
        const uint32_t d = tm + (uint32_t)((r & 1) * 256);
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
          for (int g = 0; g < 16; ++g) {
            // distinct: every repetition shifts the A view by one more row (another tap), like the conv kernels
            const uint32_t alo = a_off[g] + (p.distinct_a ? (uint32_t)(rep * (p.row_bytes >> 4)) : 0u);
            umma(d, ((uint64_t)hi_a << 32) | alo, ((uint64_t)hi_b << 32) | b_off[g], idesc, (rep | g) ? 1u : 0u);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar[r & 1])) : "memory");
      }
      __syncwarp();
    }
    for (int r = p.rounds - 2; r < p.rounds; ++r) mbar_wait(&done_bar[r & 1], (uint32_t)((r >> 1) & 1));
    const long long t1 = clock64();
    if (threadIdx.x == 0) p.clk[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* dclk;
  CK(cudaMalloc(&dclk, sms * sizeof(long long)));
  const int smem = 201 * 1024 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("# M=128 K=16 bf16, SS operands, %d CTAs (1/SM), 64 MMAs per commit\n", sms);
  printf("%-10s %4s %9s %9s %12s %12s %10s\n", "layout", "N", "distinctA", "distinctB", "clk/MMA(avg)", "ns/MMA(evt)", "TFLOP/s");
  struct Cfg { int rb, mn; const char* name; };
  const Cfg cfgs[] = {{128, 0, "K-SW128"}, {64, 0, "K-SW64"}, {32, 0, "K-SW32"}, {64, 1, "MN-SW64"}};
  for (const Cfg& c : cfgs)
    for (int da = 1; da >= 0; --da)
      for (int n : {32, 64, 96, 128, 160, 192, 224, 256}) {
        if (da == 0 && n != 64 && n != 96 && n != 192 && n != 256) continue;
        Rate p;
        p.n = n; p.row_bytes = c.rb; p.mn_major = c.mn; p.distinct_a = da; p.distinct_b = da; p.sbo_rows = 8; p.rounds = 200; p.clk = dclk;
        rate_kernel<<<sms, 128, smem>>>(p);  // warm-up
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        rate_kernel<<<sms, 128, smem>>>(p);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> h(sms);
        CK(cudaMemcpy(h.data(), dclk, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg = 0;
        for (long long v : h) avg += (double)v;
        avg /= sms;
        const double mmas = 64.0 * p.rounds;
        const double flops = 2.0 * 128 * n * 16 * mmas * sms;
        printf("%-10s %4d %9d %9d %12.1f %12.1f %10.1f\n", c.name, n, da, da, avg / mmas, ms * 1e6 / mmas, flops / (ms * 1e-3) / 1e12);
      }
  printf("\n# K-major SW128, distinct (row-shifted) A views: A-operand 8-row-group pitch (SBO) sweep -- halo tiles use the line pitch\n");
  printf("%-10s %4s %9s %12s %12s %10s\n", "layout", "N", "sbo_rows", "clk/MMA(avg)", "ns/MMA(evt)", "TFLOP/s");
  for (int n : {64, 128})
    for (int sbo : {8, 10, 16, 18, 24, 26, 34}) {
      Rate p;
      p.n = n; p.row_bytes = 128; p.mn_major = 0; p.distinct_a = 1; p.distinct_b = 1; p.sbo_rows = sbo; p.rounds = 200; p.clk = dclk;
      rate_kernel<<<sms, 128, smem>>>(p);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      rate_kernel<<<sms, 128, smem>>>(p);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::vector<long long> h(sms);
      CK(cudaMemcpy(h.data(), dclk, sms * sizeof(long long), cudaMemcpyDeviceToHost));
      double avg = 0;
      for (long long v : h) avg += (double)v;
      avg /= sms;
      const double mmas = 64.0 * p.rounds;
      printf("%-10s %4d %9d %12.1f %12.1f %10.1f\n", "K-SW128", n, sbo, avg / mmas, ms * 1e6 / mmas, 2.0 * 128 * n * 16 * mmas * sms / (ms * 1e-3) / 1e12);
    }
  return 0;
}
