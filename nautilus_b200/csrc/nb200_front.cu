// Fused fp64 front end of the cycle: one Union.sample pass
// (nautilus/bounds/union.py:305-319) + the ellipsoid half of
// NeuralBound.contains (nautilus/bounds/neural.py:117-119) + the emulator's
// input standardisation (nautilus/neural.py:115), for bounds whose mixtures
// are plain ellipsoids and that have exactly one neural bound.
//
// Per raw proposal it writes the row (8d B), one disposition byte, the
// neural-ellipsoid flag and the standardised, tf32-rounded emulator input
// (4 * round8(d) B) that k_mlp_tf32 loads straight into TMEM -- the proposal
// row is never re-read.
//
// Layout: persistent CTAs; every ellipsoid factor is staged ONCE per CTA in
// shared memory, transposed and padded to 8-row groups, so the inner loop is
//   1 strided LDS (x_j) + 1 broadcast LDS (c_j) + 4 broadcast LDS.128 (8 rows
//   of the factor) + 1 DADD + 8 DFMA,
// every output row keeping its own left-to-right FMA chain (the canonical
// order of oracle/c/nb200_oracle.c; entries above the diagonal are exact
// zeros, so including some of them changes nothing).  One shared row per
// thread: x = B y + c is formed in place, walking the row groups downwards.
#include "nb200_device.cuh"
#include "nb200_rng.cuh"

namespace nb200 {

constexpr int FRONT_THREADS = 256;

struct FrontArgs {
  int rec_off, d, d8, K, unit, k0p, stride;
  int smem_doubles;
  unsigned int lower_k;   // bit k: B_inv of mixture k is exactly lower-tri
  int lower_n;            // same for the neural bound's B_inv
  int same_k;             // mixture whose ellipsoid equals the neural bound's
                          // (-1: none)
  unsigned long long seed, offset;
  unsigned int stream_id;
  long long n;
  // optional: built-in likelihood of the rows that are still candidates
  // after the ellipsoid tests (log_l == nullptr: off); NaN for the others
  int like_id;
  const double* like_p;
  double* log_l;
};

// kept out of line so that the four likelihood bodies do not take part in
// the register allocation of the matrix-vector loops
__device__ __noinline__ double front_loglike(int like_id, const double* p,
                                             const double* x, int d) {
  return loglike_eval(like_id, p, x, d);
}

// acc[r] = sum_{j < jmax} MT[j][i0 + r] * v_j,  v_j = x[j] - (c ? c[j] : 0)
template <bool SUBTRACT>
__device__ __forceinline__ void mv8(const double* __restrict__ MT, int d8,
                                    int i0, int jmax, const double* x,
                                    const double* __restrict__ c,
                                    double (&acc)[8]) {
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.0;
  const double2* col = reinterpret_cast<const double2*>(MT + i0);
  const int ld = d8 >> 1;   // double2 per transposed row
#pragma unroll 2
  for (int j = 0; j < jmax; ++j) {
    const double v = SUBTRACT ? x[j] - c[j] : x[j];
    const double2 m0 = col[j * ld], m1 = col[j * ld + 1];
    const double2 m2 = col[j * ld + 2], m3 = col[j * ld + 3];
    acc[0] = fma(m0.x, v, acc[0]); acc[1] = fma(m0.y, v, acc[1]);
    acc[2] = fma(m1.x, v, acc[2]); acc[3] = fma(m1.y, v, acc[3]);
    acc[4] = fma(m2.x, v, acc[4]); acc[5] = fma(m2.y, v, acc[5]);
    acc[6] = fma(m3.x, v, acc[6]); acc[7] = fma(m3.y, v, acc[7]);
  }
}

// r2 = sum_i (sum_j Binv[i][j] (x_j - c_j))^2, rows in increasing order
__device__ __forceinline__ double r2_staged(const double* __restrict__ MT,
                                            int d, int d8, int lower,
                                            const double* x,
                                            const double* __restrict__ c) {
  double r2 = 0.0;
  for (int i0 = 0; i0 < d8; i0 += 8) {
    double acc[8];
    mv8<true>(MT, d8, i0, lower ? min(i0 + 8, d) : d, x, c, acc);
#pragma unroll
    for (int r = 0; r < 8; ++r) r2 = fma(acc[r], acc[r], r2);  // pad rows: 0
  }
  return r2;
}

__global__ void __launch_bounds__(FRONT_THREADS, 2)
k_front(const FrontArgs A, const int32_t* __restrict__ meta,
        const double* __restrict__ data, double* __restrict__ points,
        uint8_t* __restrict__ code, uint8_t* __restrict__ maskj,
        float* __restrict__ xs32) {
  extern __shared__ __align__(16) double sm[];
  const Rec rec{meta + A.rec_off};
  const int d = A.d, d8 = A.d8, K = A.K;
  const int mat = d * d8;
  // staged parameters
  double* BT = sm;                         // K x [d][d8]   B transposed
  double* BinvT = BT + (size_t)K * mat;    // K x [d][d8]   B_inv transposed
  double* nbT = BinvT + (size_t)K * mat;   // [d][d8]       neural B_inv^T
  double* cK = nbT + mat;                  // K x d8
  double* cN = cK + (size_t)K * d8;        // d8
  double* meanN = cN + d8;                 // d8
  double* iscaleN = meanN + d8;            // d8: 1 / scale
  double* rows = iscaleN + d8;             // FRONT_THREADS x stride
  const int32_t* nb = rec.nb(0);

  for (int e = threadIdx.x; e < (2 * K + 1) * mat; e += FRONT_THREADS) {
    const int which = e / mat, rem = e - which * mat;
    const int j = rem / d8, i = rem - j * d8;
    const double* src;
    if (which < K) src = data + rec.mix(which)[4];
    else if (which < 2 * K) src = data + rec.mix(which - K)[5];
    else src = data + nb[1];
    sm[e] = (i < d) ? src[(size_t)i * d + j] : 0.0;
  }
  for (int e = threadIdx.x; e < (K + 3) * d8; e += FRONT_THREADS) {
    const int which = e / d8, i = e - which * d8;
    double v = 0.0;
    if (i < d) {
      if (which < K) v = data[rec.mix(which)[3] + i];
      else if (which == K) v = data[nb[0] + i];
      else if (which == K + 1) v = data[nb[5] + i];
      else v = 1.0 / data[nb[6] + i];
    }
    cK[e] = v;
  }
  __syncthreads();

  const double* cdf = data + rec.off_cdf();
  const int stride = A.stride;
  double* x = rows + threadIdx.x * stride;
  const long long n_tiles = (A.n + FRONT_THREADS - 1) / FRONT_THREADS;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * FRONT_THREADS;
    const int nrows = (int)min((long long)FRONT_THREADS, A.n - base);
    const long long i = base + threadIdx.x;
    if ((int)threadIdx.x < nrows) {
      const Philox rng(A.offset + (unsigned long long)i, A.stream_id, A.seed);
      const uint4 w0 = rng.block(0);
      const double uk = u01_32(w0.x);
      int k = 0;
      while (k < K - 1 && !(uk < __ldg(cdf + k))) ++k;
      const double r = u01_32(w0.y);
      const double u = u01_53(w0.z, w0.w);
      // normals -> uniform point of the unit ball (basic.py:376-379)
      double n2 = 0.0;
      for (int j = 0; j < d; j += 4) {
        const uint4 w = rng.block(1 + (j >> 2));
        float g0, g1, g2, g3;
        normal2(w.x, w.y, g0, g1);
        normal2(w.z, w.w, g2, g3);
        const float gq[4] = {g0, g1, g2, g3};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (j + q < d) {
            const double v = (double)gq[q];
            x[j + q] = v;
            n2 = fma(v, v, n2);
          }
        }
      }
      const double scale = pow(u, 1.0 / (double)d) / sqrt(n2);
      for (int j = 0; j < d; ++j) x[j] *= scale;
      // x = B y + c in place (basic.py:380 -> :342), highest row group first
      {
        const double* Bk = BT + (size_t)k * mat;
        const double* ck = cK + k * d8;
        for (int i0 = d8 - 8; i0 >= 0; i0 -= 8) {
          double acc[8];
          mv8<false>(Bk, d8, i0, min(i0 + 8, d), x, nullptr, acc);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (i0 + q < d) x[i0 + q] = acc[q] + ck[i0 + q];
        }
      }
      uint8_t cd = NB200_CODE_IN_SHELL;
      bool in_ell = false;
      float* xrow = xs32 + i * (long long)A.k0p;
      // whitening w.r.t. the neural bound's ellipsoid: r2 and the
      // standardised, tf32-rounded emulator input row (with the constant-one
      // bias column at index d)
      auto whiten_store = [&]() -> double {
        double r2 = 0.0;
        for (int i0 = 0; i0 < A.k0p; i0 += 8) {
          uint32_t pk[8];
          if (i0 < d8) {
            double acc[8];
            mv8<true>(nbT, d8, i0, A.lower_n ? min(i0 + 8, d) : d, x, cN, acc);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              r2 = fma(acc[q], acc[q], r2);
              float v = (float)((acc[q] - meanN[i0 + q]) * iscaleN[i0 + q]);
              if (i0 + q == d) v = 1.0f;
              asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(pk[q]) : "f"(v));
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              pk[q] = (i0 + q == d) ? 0x3F800000u : 0u;
          }
          *reinterpret_cast<uint4*>(xrow + i0) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(xrow + i0 + 4) =
              make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        return r2;
      };
      double r2_nb = -1.0;
      if (A.unit && !cube_ok(x, nullptr, d)) {
        cd = NB200_CODE_CUBE_REJECT;                  // union.py:313-314
      } else {
        int nbnd = 0;                                 // union.py:316-317
        for (int kk = 0; kk < K; ++kk) {
          double r2;
          if (kk == A.same_k) {
            r2 = whiten_store();      // same ellipsoid: one pass serves both
            r2_nb = r2;
          } else {
            r2 = r2_staged(BinvT + (size_t)kk * mat, d, d8,
                           (A.lower_k >> kk) & 1, x, cK + kk * d8);
          }
          nbnd += r2 < 1.0 ? 1 : 0;
        }
        if (!(r > 1.0 - 1.0 / (double)nbnd))          // union.py:318-319
          cd = NB200_CODE_OVERLAP_REJECT;
      }
      if (cd == NB200_CODE_IN_SHELL) {
        // NeuralBound: ellipsoid test (neural.py:117)
        if (A.same_k < 0) r2_nb = whiten_store();
        in_ell = r2_nb < 1.0;
        if (!in_ell) cd = NB200_CODE_NN_REJECT;
      }
      code[i] = cd;
      maskj[i] = in_ell ? 1 : 0;
      if (A.log_l)
        A.log_l[i] = cd == NB200_CODE_IN_SHELL
                         ? front_loglike(A.like_id, A.like_p, x, d)
                         : nan("");
    }
    __syncthreads();
    store_rows(points, base, nrows, d, stride, rows);
    __syncthreads();
  }
}

// Host side: does the fast path apply, and launch.
bool front_applicable(const int32_t* meta_h, int bound, size_t* smem_out,
                      FrontArgs* args) {
  const Rec rec = record(meta_h, bound);
  if (rec.kind() != 1 || rec.J() != 1) return false;
  const int32_t* nb = rec.nb(0);
  if (nb[3] <= 0 || nb[10] < 0 || nb[11] <= 0) return false;
  const int d = rec.d(), K = rec.K();
  if (K > 32) return false;
  unsigned int lower_k = 0;
  for (int k = 0; k < K; ++k) {
    if (rec.mix(k)[1] != 0 || rec.mix(k)[0] != d) return false;
    if (rec.mix(k)[6]) lower_k |= 1u << k;
  }
  const int d8 = (d + 7) / 8 * 8;
  const size_t doubles = (size_t)(2 * K + 1) * d * d8 + (size_t)(K + 3) * d8 +
                         (size_t)FRONT_THREADS * (d | 1);
  if (doubles * 8 > 110 * 1024) return false;   // two CTAs per SM
  if (smem_out) *smem_out = doubles * 8;
  if (args) {
    args->rec_off = (int)(rec.r - meta_h);
    args->d = d; args->d8 = d8; args->K = K; args->unit = rec.unit();
    args->k0p = (d + 1 + 7) / 8 * 8; args->stride = d | 1;
    args->smem_doubles = (int)doubles;
    args->lower_k = lower_k; args->lower_n = nb[2];
    args->same_k = rec.r[10] - 1;
  }
  return true;
}

int launch_front(const int32_t* meta_h, const int32_t* meta_d,
                 const double* data_d, int bound, int64_t n, uint64_t seed,
                 uint64_t offset, uint32_t stream_id, double* points,
                 uint8_t* code, uint8_t* maskj, float* xs32, int like_id,
                 const double* like_p, double* log_l, cudaStream_t st) {
  FrontArgs A;
  size_t smem = 0;
  NB_CHECK(front_applicable(meta_h, bound, &smem, &A), "front kernel n/a");
  A.seed = seed; A.offset = offset; A.stream_id = stream_id; A.n = n;
  A.like_id = like_id; A.like_p = like_p; A.log_l = log_l;
  NB_CUDA(cudaFuncSetAttribute(k_front,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = (n + FRONT_THREADS - 1) / FRONT_THREADS;
  int64_t grid = 2 * (int64_t)sms;
  if (grid > n_tiles) grid = n_tiles;
  ProfScope prof(ST_FUSED, st);
  k_front<<<(unsigned)grid, FRONT_THREADS, smem, st>>>(A, meta_d, data_d,
                                                       points, code, maskj,
                                                       xs32);
  NB_LAUNCH_OK();
  return 0;
}

}  // namespace nb200
