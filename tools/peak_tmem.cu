// Measures the TMEM read / write bandwidth of one SM (tcgen05.ld / .st,
// 32x32b.x32) on this GPU: the ceiling of the emulator kernel's epilogues,
// which read every accumulator column once and write the hidden ones back.
// W warps per CTA (warp w owns TMEM lanes 32 (w % 4) ...), one CTA per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
#define R32(v) "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),"=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31])
#define W32(v) "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]),"r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31])
#define REGS "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}"
#define REGS1 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32}"

template <int MODE>   // 0: loads, 1: stores, 2: load + store (the epilogue mix)
__global__ void k(int iters, uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (((uint32_t)(warp & 3) * 32) << 16);
  uint32_t v[32];
  for (int q = 0; q < 32; ++q) v[q] = threadIdx.x + q;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < 512; c += 32) {
      if (MODE != 1) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " REGS ", [%32];"
                     : R32(v) : "r"(base + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= v[0] ^ v[13] ^ v[31];
      }
      if (MODE != 0) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], " REGS1 ";"
                     ::"r"(base + c), W32(v) : "memory");
      }
    }
    if (MODE != 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(slot), "r"(512) : "memory");
}

template <int MODE>
double run(int sms, int warps, uint32_t* out, double mhz) {
  const int iters = 2000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms, warps * 32>>>(10, out); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); k<MODE><<<sms, warps * 32>>>(iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  // bytes moved per SM: warps x 16 chunks x 4 KB per iteration (x2 in mode 2)
  const double bytes = (double)warps * 16 * 4096 * iters * (MODE == 2 ? 2 : 1);
  return bytes / (best * 1e-3) / (mhz * 1e6);      // bytes per clock per SM
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  uint32_t* out; cudaMalloc(&out, sizeof(uint32_t) * p.multiProcessorCount * 1024);
  printf("{\"gpu\": \"%s\", \"sm_mhz_max\": %.0f", p.name, mhz);
  for (int w : {4, 8, 16}) {
    printf(", \"ld_B_per_clk_%dw\": %.1f", w, run<0>(p.multiProcessorCount, w, out, mhz));
    printf(", \"st_B_per_clk_%dw\": %.1f", w, run<1>(p.multiProcessorCount, w, out, mhz));
    printf(", \"ldst_B_per_clk_%dw\": %.1f", w, run<2>(p.multiProcessorCount, w, out, mhz));
  }
  printf("}\n");
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
  return 0;
}
