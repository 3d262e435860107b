// Measures the non-tensor FP64 FMA peak: 8
// independent DFMA chains per thread, enough warps to fill every SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3,
         a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c);
    a3 = fma(a3, b, c); a4 = fma(a4, b, c); a5 = fma(a5, b, c);
    a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] =
      a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, threads>>>(out, 1024); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0); k<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double flops = 2.0 * 8 * iters * (double)blocks * threads;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.2f, "
         "\"ms\": %.3f}\n", p.name, p.multiProcessorCount,
         flops / (best * 1e-3) / 1e12, best);
  return 0;
}
