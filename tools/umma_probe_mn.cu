// Probe: MN-major A-operand descriptors (SW128 / SW64) over a voxel-row tile: rows = K (voxels), channels = M.
// Question (needed for a halo-resident weight-gradient kernel): may the start address be any ROW, may the 8-row-group
// stride (SBO) be arbitrary, and may the M-block stride (LBO) be ONE ROW, i.e. may the M blocks be overlapping
// row-shifted views of the same tile (block j = the tile shifted by j*lbo rows)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe_mn tools/umma_probe_mn.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

struct Probe {
  CUtensorMap amap, bmap;
  float* out;       // [128][32]
  int row_bytes;    // 128 (SW128, 64 ch) or 64 (SW64, 32 ch)
  int start_rows;   // descriptor start offset in rows
  int sbo_rows;     // stride between 8-row groups, in rows
  int lbo_rows;     // stride between M blocks (row_bytes/2 channels each), in rows
  int arows;        // rows loaded for A
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ Probe p) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = sm;
  uint8_t* sb = sm + 48 * 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&full_bar, 1); mbar_init(&done_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  const int kel = p.row_bytes / 2;  // K elements per row
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, p.arows * p.row_bytes + 32 * 64);
    for (int r0 = 0; r0 < p.arows; r0 += 128) tma_2d(sa + r0 * p.row_bytes, &p.amap, &full_bar, 0, r0);
    tma_2d(sb, &p.bmap, &full_bar, 0, 0);
    mbar_wait(&full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t layout = p.row_bytes == 128 ? 2ull : 4ull;
    // A MN-major (bit 15), B K-major, N = 32, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    for (int k = 0; k < 2; ++k) {  // K total = 32 voxels = 2 MMAs of 16 (two 8-row groups each)
      const uint32_t a_addr = smem_u32(sa) + (p.start_rows + 2 * k * p.sbo_rows) * p.row_bytes;
      const uint32_t b_addr = smem_u32(sb) + k * 32;
      uint64_t da = 0, db = 0;
      da |= (uint64_t)((a_addr & 0x3FFFFu) >> 4);
      da |= (uint64_t)(((uint32_t)(p.lbo_rows * p.row_bytes) >> 4) & 0x3FFFu) << 16;
      da |= (uint64_t)(((uint32_t)(p.sbo_rows * p.row_bytes) >> 4) & 0x3FFFu) << 32;
      da |= 1ull << 46;
      da |= layout << 61;
      db |= (uint64_t)((b_addr & 0x3FFFFu) >> 4);
      db |= (uint64_t)(((uint32_t)(8 * 64) >> 4) & 0x3FFFu) << 32;
      db |= 1ull << 46;
      db |= 4ull << 61;  // B: [32 n][32 k] bf16, 64-byte rows, SW64
      const uint32_t acc = k > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
  }
  mbar_wait(&done_bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(tm + ((uint32_t)(warp * 32) << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) p.out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32u) : "memory");
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncFn enc = (EncFn)fp;
  const int AROWS = 384;
  for (int row_bytes : {64, 128}) {
    const int cb = row_bytes / 2;       // channels per row = M-block width
    const int nblk = 128 / cb;
    std::vector<__nv_bfloat16> hA(AROWS * cb), hB(32 * 32);
    std::vector<float> fA(AROWS * cb), fB(32 * 32);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; hA[i] = __float2bfloat16(v); fA[i] = v; }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 13 - 6) / 4.f; hB[i] = __float2bfloat16(v); fB[i] = v; }
    __nv_bfloat16 *dA, *dB; float* dO;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dO, 128 * 32 * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    Probe p;
    cuuint32_t es[2] = {1, 1};
    CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    { cuuint64_t dims[2] = {(cuuint64_t)cb, AROWS}; cuuint64_t st[1] = {(cuuint64_t)row_bytes}; cuuint32_t box[2] = {(cuuint32_t)cb, 128};
      CUresult r = enc(&p.amap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode A failed %d\n", r); return 1; } }
    { cuuint64_t dims[2] = {32, 32}; cuuint64_t st[1] = {64}; cuuint32_t box[2] = {32, 32};
      CUresult r = enc(&p.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode B failed %d\n", r); return 1; } }
    p.out = dO; p.row_bytes = row_bytes; p.arows = AROWS;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const int starts[] = {0, 8, 1, 2, 3, 5, 9, 13};
    const int sbos[] = {8, 10, 18, 34};
    const int lbos[] = {128, 1, 2, 3, 18, 34};
    for (int lbo : lbos) for (int sbo : sbos) for (int st : starts) {
      if (st + (nblk - 1) * lbo + 3 * sbo + 8 > AROWS) continue;
      p.start_rows = st; p.sbo_rows = sbo; p.lbo_rows = lbo;
      CK(cudaMemset(dO, 0, 128 * 32 * 4));
      probe_kernel<<<1, 128, 64 * 1024>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("row_bytes %d lbo %d sbo %d start %d: KERNEL ERROR %s\n", row_bytes, lbo, sbo, st, cudaGetErrorString(e)); return 1; }
      std::vector<float> out(128 * 32);
      CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int m = 0; m < 128; ++m) {
        const int blk = m / cb, c = m % cb;
        for (int n = 0; n < 32; ++n) {
          double ref = 0;
          for (int k = 0; k < 32; ++k) {
            const int arow = st + blk * lbo + (k / 8) * sbo + (k % 8);
            ref += (double)fA[arow * cb + c] * fB[n * 32 + k];
          }
          maxerr = fmax(maxerr, fabs(ref - out[m * 32 + n]));
        }
      }
      printf("MN-major row_bytes %3d  lbo_rows %3d  sbo_rows %2d  start_row %2d  -> max err %.4f  %s\n", row_bytes, lbo, sbo, st, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
  }
  return 0;
}
