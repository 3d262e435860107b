// Measures the tcgen05.mma peak of this GPU in the shapes the emulator kernels
// use: cta_group::1, M = 128, kind::tf32 (K = 8 per instruction) and
// kind::f16 (K = 16), A operand from TMEM (".ts", k_mlp_tf32 / k_mlp_fit_tc)
// or from shared memory (".ss", the trainer's gradient products), B from a
// K-major no-swizzle shared-memory tile, N in {64, 128, 256}.  One CTA per
// SM, one issuing thread, every instruction accumulating into the same TMEM
// tile (as a GEMM main loop does).  The result is the denominator of the
// `emulator_tensor` fraction in bench.py: the bf16 cuBLAS figure of
// MEASURED_PEAKS.json is a cta_group::2 / N=256 number and, for tf32, the
// wrong data type.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peak_umma tools/peak_umma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../nautilus_b200/csrc/nb200_tc.cuh"

using namespace nb200;

__device__ __forceinline__ void mma_ss(bool f16, uint32_t d, uint64_t a,
                                       uint64_t b, uint32_t id, uint32_t acc) {
  if (f16)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}

// MODE bit 0: kind::f16 instead of kind::tf32; bit 1: A from shared memory
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int iters, int n, uint32_t* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  constexpr bool F16 = MODE & 1, SS = MODE & 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  // operands: small non-zero numbers (both data types read them as finite)
  for (int e = tid; e < 48 * 1024 / 4; e += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[e] = F16 ? 0x2C002E00u + (e & 0xFF)
                                               : 0x3C000000u + ((e & 0xFF) << 13);
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  // A in TMEM: columns 256.. of this warp's 32 lanes
  {
    uint32_t v[32];
    for (int q = 0; q < 32; ++q)
      v[q] = F16 ? 0x2C002E00u + q : 0x3C000000u + ((uint32_t)(q + tid) << 13);
    tmem_st32(tmem + (((uint32_t)warp * 32) << 16) + 256, v);
    tmem_st32(tmem + (((uint32_t)warp * 32) << 16) + 288, v);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    // B tile: n rows x 32 bytes of K, core matrices of 8 rows x 16 B; A tile
    // (ss): 128 rows x 32 bytes, after it
    const uint32_t b_addr = smem_u32(smem);
    const uint32_t a_addr = b_addr + 16 * 1024;
    const uint64_t bd = smem_desc(b_addr, 128u, 256u);
    const uint64_t ad = smem_desc(a_addr, 128u, 256u);
    const uint32_t id = F16 ? idesc_f16(n) : idesc_tf32(n);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const uint32_t acc = (it | s) ? 1u : 0u;
        if (SS) {
          mma_ss(F16, tmem, ad, bd, id, acc);
        } else if (F16) {
          mma_f16_ts(tmem, tmem + 256 + 8 * s, bd, id, acc);
        } else {
          mma_tf32_ts(tmem, tmem + 256 + 8 * s, bd, id, acc);
        }
      }
    }
    mma_commit(&bar);
  }
  {
    // (not mbar_wait: a launch here runs for milliseconds)
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      if (!done) __nanosleep(500);
    }
  }
  tc_fence_after();
  uint32_t v[8];
  tmem_ld8(tmem + (((uint32_t)warp * 32) << 16), v);
  tmem_wait_ld();
  out[blockIdx.x * blockDim.x + tid] = v[0] ^ v[7];
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(slot), "r"(512) : "memory");
}

template <int MODE>
double run(int sms, int n, uint32_t* out) {
  const int smem = 48 * 1024;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 20000;
  k<MODE><<<sms, 128, smem>>>(200, n, out);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE><<<sms, 128, smem>>>(iters, n, out);
  cudaEventRecord(b);
  if (cudaEventSynchronize(b) != cudaSuccess) return -1.0;
  float ms = 0.f; cudaEventElapsedTime(&ms, a, b);
  const double kk = (MODE & 1) ? 16.0 : 8.0;
  const double flops = 2.0 * 128.0 * n * kk * 8.0 * iters * sms;
  return flops / (ms * 1e-3) / 1e12;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  uint32_t* out; cudaMalloc(&out, sizeof(uint32_t) * p.multiProcessorCount * 128);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"unit\": \"TFLOP/s\", \"shape\": "
         "\"cta_group::1, M=128, one CTA per SM, 160000 MMAs per SM\"",
         p.name, p.multiProcessorCount);
  for (int n : {64, 128, 256}) {
    printf(", \"tf32_ts_n%d\": %.1f", n, run<0>(p.multiProcessorCount, n, out));
    printf(", \"f16_ts_n%d\": %.1f", n, run<1>(p.multiProcessorCount, n, out));
    printf(", \"tf32_ss_n%d\": %.1f", n, run<2>(p.multiProcessorCount, n, out));
    printf(", \"f16_ss_n%d\": %.1f", n, run<3>(p.multiProcessorCount, n, out));
  }
  printf("}\n");
  if (cudaDeviceSynchronize() != cudaSuccess) {
    printf("error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  return 0;
}
