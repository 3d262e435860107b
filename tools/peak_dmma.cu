// Measures the FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) peak on this GPU,
// next to the DFMA peak of tools/peak_fp64.cu: decides whether the fp64
// matrix-vector products of k_front should move to DMMA.  W independent
// accumulator tiles per warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int W>
__global__ void k(double* out, int iters) {
  double c[W][2];
  for (int w = 0; w < W; ++w) { c[w][0] = threadIdx.x * 1e-9 + w; c[w][1] = w; }
  const double a = 1.0000001, b = 0.25;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int w = 0; w < W; ++w)
      asm volatile(
          "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 "
          "{%0,%1}, {%2}, {%3}, {%0,%1};"
          : "+d"(c[w][0]), "+d"(c[w][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int w = 0; w < W; ++w) s += c[w][0] + c[w][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int W>
double run(int blocks, int threads, double* out) {
  const int iters = 1 << 14;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<W><<<blocks, threads>>>(out, 256); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0); k<W><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  // one m8n8k4 = 256 FMA = 512 flop per warp
  const double flops = 512.0 * W * iters * (double)blocks * (threads / 32);
  return flops / (best * 1e-3) / 1e12;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 8 * 256);
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, p.multiProcessorCount);
  printf(", \"dmma_tflops_w1_16warps\": %.2f", run<1>(p.multiProcessorCount * 2, 256, out));
  printf(", \"dmma_tflops_w4_16warps\": %.2f", run<4>(p.multiProcessorCount * 2, 256, out));
  printf(", \"dmma_tflops_w8_16warps\": %.2f", run<8>(p.multiProcessorCount * 2, 256, out));
  printf(", \"dmma_tflops_w8_8warps\": %.2f", run<8>(p.multiProcessorCount * 2, 128, out));
  printf(", \"dmma_tflops_w8_64warps\": %.2f", run<8>(p.multiProcessorCount * 8, 256, out));
  printf("}\n");
  return 0;
}
