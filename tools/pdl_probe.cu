// How much of the gap between two dependent launches in one stream does programmatic dependent launch hide on this part?
// A chain of N short kernels (each ~5 us of work on 148 CTAs), launched (a) normally, (b) with the programmatic stream
// serialization attribute + griddepcontrol.wait at the top, (c) as (b) + griddepcontrol.launch_dependents at the top.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pdl_probe tools/pdl_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void k(float* p, int iters) {
  if (MODE >= 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (MODE >= 1) asm volatile("griddepcontrol.wait;" ::: "memory");
  float v = p[blockIdx.x * blockDim.x + threadIdx.x];
  for (int i = 0; i < iters; ++i) v = fmaf(v, 1.0001f, 0.5f);
  p[blockIdx.x * blockDim.x + threadIdx.x] = v;
}

template <int MODE>
static float run(float* d, int n, int iters, cudaStream_t s) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = MODE >= 1 ? 1 : 0;
  for (int w = 0; w < 2; ++w) {
    cudaEventRecord(e0, s);
    for (int i = 0; i < n; ++i) cudaLaunchKernelEx(&cfg, k<MODE>, d, iters);
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  float* d; cudaMalloc(&d, 148 * 256 * 4); cudaMemset(d, 0, 148 * 256 * 4);
  cudaStream_t s; cudaStreamCreate(&s);
  const int n = 400;
  for (int iters : {200, 2000, 20000}) {
    float a = run<0>(d, n, iters, s), b = run<1>(d, n, iters, s), c = run<2>(d, n, iters, s);
    printf("iters %6d: plain %.2f us/launch, pdl(wait) %.2f, pdl(wait+trigger) %.2f\n", iters, a * 1000 / n, b * 1000 / n, c * 1000 / n);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
