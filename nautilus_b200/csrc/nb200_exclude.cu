// Later-bound exclusion as ONE grouped pass (HOT LOOP D of the reference,
// nautilus/sampler.py:796-801: a proposal of shell i is dropped if ANY later
// bound contains it; 70 % of the reference's sampling-phase time).
//
// The reference calls bound.contains(points) once per later bound
// (bounds/nautilus.py:146-169 -> union.py:269-289 + neural.py:99-126).  Here:
//   1. the proposals that are still in the shell are compacted to a candidate
//      list (the exclusion never looks at the other ~85 % of the raw batch);
//   2. k_excl_prep: ONE launch tests every candidate against every later
//      bound in fp64 -- union membership (all mixtures + unit cube), then each
//      neural bound's ellipsoid -- with the candidate's row loaded into shared
//      memory once for all bounds.  A (candidate, later bound, neural bound)
//      triple that still needs the emulator's verdict is appended, already
//      whitened, standardised and tf32-rounded, to the segment of its
//      (bound, neural bound) pair;
//   3. k_mlp_tf32<GROUPED>: ONE launch runs all segments through the tcgen05
//      emulator, each with its pair's weights streamed out of L2 and its
//      pair's threshold, and marks the candidates that a later bound contains;
//   4. k_excl_apply writes NB200_CODE_EXCLUDED back into the dispositions.
// The launch count does not depend on the number of later bounds.  The fp64
// arithmetic is that of k_union_count / k_neural_prep / k_standardise_tf32
// (same device functions, same expressions), so the result is bit-identical
// to the per-bound loop it replaces (tests/test_gpu_exclusion.py).
#include "nb200_device.cuh"
#include "nb200_tc.cuh"

namespace nb200 {

constexpr int EX_THREADS = 128;

// pair tables: pair_base[l] = first pair of later bound l; one pair per
// (later bound, neural bound), with or without an emulator
__global__ void k_excl_pairs(const int32_t* __restrict__ meta, int first_later,
                             int n_later, int f16,
                             PairRec* __restrict__ pairs,
                             int* __restrict__ pair_base) {
  // one thread per later bound; the prefix over the J's by thread 0
  extern __shared__ int sJ[];
  for (int l = threadIdx.x; l < n_later; l += blockDim.x) {
    const Rec rec = record(meta, first_later + l);
    sJ[l] = rec.kind() == 1 ? rec.J() : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int p = 0;
    for (int l = 0; l < n_later; ++l) {
      const int J = sJ[l];
      sJ[l] = p;
      pair_base[l] = p;
      p += J;
    }
  }
  __syncthreads();
  for (int l = threadIdx.x; l < n_later; l += blockDim.x) {
    const Rec rec = record(meta, first_later + l);
    const int J = rec.kind() == 1 ? rec.J() : 0;
    for (int j = 0; j < J; ++j) {
      const int32_t* nb = rec.nb(j);
      PairRec pr;
      pr.rec_off = (int)(rec.r - meta);
      pr.j = j;
      // (the fp16 header sits 32 ints behind the tf32 one; its word 20 is
      // the position of the fp16 blob, _pack.py:pack_record)
      pr.blob_off = nb[3] <= 0 ? -1 : f16 ? rec.r[nb[11] + 32 + 20] : nb[10];
      pr.thr_off = nb[3] > 0 ? nb[7] : -1;
      pairs[sJ[l] + j] = pr;
    }
  }
}

// matvec_rows (nb200_device.cuh) for a factor staged in SHARED memory: same
// left-to-right FMA chains, plain loads
template <typename F>
__device__ __forceinline__ void matvec_rows_sm(const double* M, int de,
                                               int lower, const double* s,
                                               F&& f) {
  int i = 0;
  for (; i + 4 <= de; i += 4) {
    const double* m0 = M + (size_t)i * de;
    const double* m1 = m0 + de;
    const double* m2 = m1 + de;
    const double* m3 = m2 + de;
    const int jmax = lower ? i + 4 : de;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 4
    for (int j = 0; j < jmax; ++j) {
      const double sj = s[j];
      a0 = fma(m0[j], sj, a0);
      a1 = fma(m1[j], sj, a1);
      a2 = fma(m2[j], sj, a2);
      a3 = fma(m3[j], sj, a3);
    }
    f(i, a0); f(i + 1, a1); f(i + 2, a2); f(i + 3, a3);
  }
  for (; i < de; ++i) {
    const double* m0 = M + (size_t)i * de;
    const int jmax = lower ? i + 1 : de;
    double a0 = 0.0;
    for (int j = 0; j < jmax; ++j) a0 = fma(m0[j], s[j], a0);
    f(i, a0);
  }
}

// append one (candidate, pair) row to the pair's segment
__device__ __forceinline__ void excl_append(const float* f, int d, int k0p,
                                            int f16, int p, long long c,
                                            unsigned int* seg_count,
                                            float* xs, unsigned int* cid,
                                            long long seg_stride) {
  const unsigned int slot = atomicAdd(seg_count + p, 1u);
  const long long row = (long long)p * seg_stride + slot;
  if (f16) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(xs) +
                    row * (long long)(k0p >> 1);
    for (int q = 0; q < k0p; q += 2) {
      const float v0 = q < d ? f[q] : (q == d ? 1.0f : 0.0f);
      const float v1 = q + 1 < d ? f[q + 1] : (q + 1 == d ? 1.0f : 0.0f);
      dst[q >> 1] = pack_f16x2(v0, v1);
    }
  } else {
    float* dst = xs + row * (long long)k0p;
    for (int q = 0; q < k0p; ++q) {
      const float v = q < d ? f[q] : (q == d ? 1.0f : 0.0f);
      uint32_t rr;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rr) : "f"(v));
      dst[q] = __uint_as_float(rr);
    }
  }
  cid[row] = (unsigned int)(c);
}

__global__ void __launch_bounds__(EX_THREADS)
k_excl_prep(const int32_t* __restrict__ meta, const double* __restrict__ data,
            int first_later, int n_later, int d, int k0p, int f16,
            const double* __restrict__ points,
            const unsigned long long* __restrict__ cand_idx,
            const unsigned long long* __restrict__ n_cand, long long chunk_lo,
            long long chunk_cap, const int* __restrict__ pair_base,
            unsigned int* __restrict__ seg_count, float* __restrict__ xs,
            unsigned int* __restrict__ cid, long long seg_stride,
            uint8_t* __restrict__ excl) {
  extern __shared__ double sm[];
  const int stride = row_stride(d);
  double* rowX = sm;                                      // the candidate
  double* rowS = rowX + (size_t)EX_THREADS * stride;      // x - c (scratch)
  // parameters of the later bound at hand, staged once per block (the usual
  // bound: one plain ellipsoid + one neural bound): the factors come out of
  // shared memory instead of L2 (47 bounds x 2 factors do not fit L1)
  double* stM = rowS + (size_t)EX_THREADS * stride;       // d*d  mixture B_inv
  double* stN = stM + d * d;                              // d*d  neural B_inv
  double* stV = stN + d * d;        // 4 d: c (mixture), c (neural), mean, 1/scale
  float* rowF = reinterpret_cast<float*>(stV + 4 * d);
  const long long m = (long long)*n_cand;
  const long long hi = min(m, chunk_lo + chunk_cap);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long t0 = chunk_lo + (long long)blockIdx.x * EX_THREADS; t0 < hi;
       t0 += (long long)gridDim.x * EX_THREADS) {
    const int nrows = (int)min((long long)EX_THREADS, hi - t0);
    // gather the candidates' rows: one warp per row, coalesced
    for (int r = warp; r < nrows; r += EX_THREADS / 32) {
      const double* src = points + cand_idx[t0 + r] * (unsigned long long)d;
      for (int j = lane; j < d; j += 32) rowX[r * stride + j] = src[j];
    }
    __syncthreads();
    const bool valid = (int)threadIdx.x < nrows;
    const long long c = t0 + threadIdx.x;
    const double* x = rowX + threadIdx.x * stride;
    double* s = rowS + threadIdx.x * stride;
    float* f = rowF + threadIdx.x * k0p;
    bool out = false;
    for (int l = 0; l < n_later; ++l) {
      const Rec rec = record(meta, first_later + l);
      const bool plain = rec.kind() == 1 && rec.K() == 1 && rec.J() == 1 &&
                         rec.mix(0)[1] == 0 && rec.mix(0)[0] == d;
      if (plain) {
        // ---- staged: one ellipsoid, one neural bound (block-uniform) --------
        const int32_t* mix = rec.mix(0);
        const int32_t* nb = rec.nb(0);
        const bool same = rec.r[10] - 1 == 0;    // neural ellipsoid == mixture's
        const bool has_emu = nb[3] > 0;
        __syncthreads();                         // the previous bound is done
        for (int e = threadIdx.x; e < d * d; e += EX_THREADS) {
          stM[e] = __ldg(data + mix[5] + e);
          if (!same) stN[e] = __ldg(data + nb[1] + e);
        }
        for (int e = threadIdx.x; e < d; e += EX_THREADS) {
          stV[e] = __ldg(data + mix[3] + e);
          stV[d + e] = __ldg(data + nb[0] + e);
          if (has_emu) {
            stV[2 * d + e] = __ldg(data + nb[5] + e);
            stV[3 * d + e] = 1.0 / __ldg(data + nb[6] + e);
          }
        }
        __syncthreads();
        if (!valid || out) continue;
        // Union.contains (union.py:285-289): the mixture's ellipsoid + cube
        for (int q = 0; q < d; ++q) s[q] = x[q] - stV[q];
        double r2 = 0.0;
        matvec_rows_sm(stM, d, mix[6], s, [&](int i, double v) {
          r2 = fma(v, v, r2);
          if (same && has_emu)
            f[i] = (float)((v - stV[2 * d + i]) * stV[3 * d + i]);
        });
        bool in = r2 < 1.0;
        if (in && rec.unit()) in = cube_ok(x, nullptr, d);
        if (!in) continue;
        // NeuralBound.contains (neural.py:115-119)
        if (!same) {
          for (int q = 0; q < d; ++q) s[q] = x[q] - stV[d + q];
          r2 = 0.0;
          matvec_rows_sm(stN, d, nb[2], s, [&](int i, double v) {
            r2 = fma(v, v, r2);
            if (has_emu) f[i] = (float)((v - stV[2 * d + i]) * stV[3 * d + i]);
          });
          if (!(r2 < 1.0)) continue;
        }
        if (!has_emu) { out = true; continue; }
        excl_append(f, d, k0p, f16, pair_base[l], c, seg_count, xs, cid,
                    seg_stride);
        continue;
      }
      // ---- any other bound: parameters read through the cache --------------
      if (!valid || out) continue;
      // Union.contains (union.py:285-289); a cube record is the unit cube
      bool in;
      if (rec.kind() == 0) {
        in = cube_ok(x, nullptr, d);
      } else {
        in = union_count(rec, data, x, s) > 0;
        if (in && rec.unit()) in = cube_ok(x, nullptr, d);
      }
      if (!in) continue;
      const int J = rec.kind() == 1 ? rec.J() : 0;
      if (J == 0) { out = true; continue; }
      for (int j = 0; j < J; ++j) {
        const int32_t* nb = rec.nb(j);
        const bool has_emu = nb[3] > 0;
        const double* cN = data + nb[0];
        const double* mean = data + nb[5];
        const double* scale = data + nb[6];
        // NeuralBound.contains (neural.py:115-119): ellipsoid test and the
        // whitened coordinates, standardised for the emulator
        // (nautilus/neural.py:115) exactly as k_standardise_tf32 does
        for (int q = 0; q < d; ++q) s[q] = x[q] - __ldg(cN + q);
        double r2 = 0.0;
        matvec_rows(data + nb[1], d, nb[2], s, [&](int i, double v) {
          r2 = fma(v, v, r2);
          if (has_emu)
            f[i] = (float)((v - __ldg(mean + i)) * (1.0 / __ldg(scale + i)));
        });
        if (!(r2 < 1.0)) continue;
        if (!has_emu) { out = true; break; }
        excl_append(f, d, k0p, f16, pair_base[l] + j, c, seg_count, xs, cid,
                    seg_stride);
      }
    }
    if (valid && out) excl[c] = 1;
    __syncthreads();
  }
}

__global__ void k_excl_apply(const unsigned long long* __restrict__ cand_idx,
                             const unsigned long long* __restrict__ n_cand,
                             const uint8_t* __restrict__ excl,
                             uint8_t* __restrict__ code) {
  const long long m = (long long)*n_cand;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < m;
       c += (long long)gridDim.x * blockDim.x)
    if (excl[c]) code[cand_idx[c]] = NB200_CODE_EXCLUDED;
}

// ---- host wrappers ----------------------------------------------------------

int launch_excl_pairs(const int32_t* meta_d, int first_later, int n_later,
                      int f16, PairRec* pairs, int* pair_base,
                      cudaStream_t st) {
  k_excl_pairs<<<1, 128, sizeof(int) * (size_t)(n_later > 0 ? n_later : 1),
                 st>>>(meta_d, first_later, n_later, f16, pairs, pair_base);
  NB_LAUNCH_OK();
  return 0;
}

size_t excl_prep_smem(int d, int k0p) {
  return (size_t)EX_THREADS * (2 * (size_t)(d | 1) * sizeof(double) +
                               (size_t)k0p * sizeof(float)) +
         (2 * (size_t)d * d + 4 * (size_t)d) * sizeof(double);
}

int launch_excl_prep(const int32_t* meta_d, const double* data_d,
                     int first_later, int n_later, int d, int k0p, int f16,
                     const double* points, const unsigned long long* cand_idx,
                     const unsigned long long* n_cand, long long chunk_lo,
                     long long chunk_cap, const int* pair_base,
                     unsigned int* seg_count, float* xs, unsigned int* cid,
                     long long seg_stride, uint8_t* excl, cudaStream_t st) {
  const size_t smem = excl_prep_smem(d, k0p);
  NB_CHECK(smem <= 227 * 1024, "exclusion prep: rows do not fit shared memory");
  NB_CUDA(cudaFuncSetAttribute(k_excl_prep,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long grid = (chunk_cap + EX_THREADS - 1) / EX_THREADS;
  const long long resident = (long long)sms * (smem * 2 <= 220 * 1024 ? 2 : 1);
  if (grid > 4 * resident) grid = 4 * resident;
  if (grid < 1) grid = 1;
  k_excl_prep<<<(unsigned)grid, EX_THREADS, smem, st>>>(
      meta_d, data_d, first_later, n_later, d, k0p, f16, points, cand_idx,
      n_cand, chunk_lo, chunk_cap, pair_base, seg_count, xs, cid, seg_stride,
      excl);
  NB_LAUNCH_OK();
  return 0;
}

int launch_excl_apply(const unsigned long long* cand_idx,
                      const unsigned long long* n_cand, const uint8_t* excl,
                      uint8_t* code, int64_t n, cudaStream_t st) {
  long long grid = (n + 255) / 256;
  if (grid > 2048) grid = 2048;
  if (grid < 1) grid = 1;
  k_excl_apply<<<(unsigned)grid, 256, 0, st>>>(cand_idx, n_cand, excl, code);
  NB_LAUNCH_OK();
  return 0;
}

}  // namespace nb200
