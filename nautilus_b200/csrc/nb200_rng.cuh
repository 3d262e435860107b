// Philox4x32-10 counter RNG (Salmon et al., SC'11) and the conversions the
// proposal kernels use.  Mirrored bit-for-bit by oracle/philox.py.
//
// Replaces the reference's serial PCG64 stream (nautilus/sampler.py:305,
// shared with every bound at :1002,1031): counter = (global proposal index,
// block, stream id), key = sampler seed, so a proposal's randoms do not depend
// on how the batch is sharded over threads, launches or GPUs.
#pragma once
#include <stdint.h>

namespace nb200 {

struct Philox {
  uint32_t c0, c1, c3;  // proposal index (lo, hi) and stream id
  uint32_t k0, k1;

  __device__ __forceinline__ Philox(uint64_t idx, uint32_t stream,
                                    uint64_t seed)
      : c0((uint32_t)idx), c1((uint32_t)(idx >> 32)), c3(stream),
        k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}

  __device__ __forceinline__ uint4 block(uint32_t b) const {
    uint32_t x0 = c0, x1 = c1, x2 = b, x3 = c3, ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
      const uint32_t n0 = hi1 ^ x1 ^ ka, n2 = hi0 ^ x3 ^ kb;
      x0 = n0; x1 = lo1; x2 = n2; x3 = lo0;
      ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    return make_uint4(x0, x1, x2, x3);
  }
};

// (w + 0.5) * 2^-32 in fp64: uniform on (0,1), exact.
__device__ __forceinline__ double u01_32(uint32_t w) {
  return ((double)w + 0.5) * 2.3283064365386963e-10;
}
// 53-bit uniform on [0,1): 27 + 26 bits, exact.
__device__ __forceinline__ double u01_53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) *
         1.1102230246251565e-16;
}
// Two standard normals by Box-Muller in fp32 (radius <= 5.7 sigma) on the
// special-function unit (MUFU lg2 / sqrt / sin / cos).  The normals only fix a
// direction on the sphere, so ~1e-6 absolute accuracy is ample; everything
// downstream is fp64.  The radial uniform uses 23 bits: (k + 0.5) * 2^-23 is
// exact in fp32 for every k < 2^23 and therefore strictly inside (0, 1) -- with
// 24 bits the largest value rounded to 1.0f, i.e. a radius of exactly 0, and a
// d <= 2 ellipsoid part would have divided by |z| = 0 once in 2^24 draws.
__device__ __forceinline__ void normal2(uint32_t a, uint32_t b, float& n0,
                                        float& n1) {
  const float u1 = ((float)(a >> 9) + 0.5f) * 1.1920928955078125e-7f;
  const float ang =
      ((float)(b >> 8) * 5.9604644775390625e-8f - 0.5f) * 6.283185307179586f;
  // (sqrt.approx: one MUFU op, ~1 ulp; sqrtf is a MUFU.RSQ plus a Newton
  // step and a range fix-up)
  float rr;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rr) : "f"(-2.0f * __logf(u1)));
  n0 = rr * __cosf(ang);
  n1 = rr * __sinf(ang);
}

}  // namespace nb200
