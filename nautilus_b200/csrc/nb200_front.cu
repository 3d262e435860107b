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
#include <stdlib.h>

#include "nb200_device.cuh"
#include "nb200_rng.cuh"
#include "nb200_tc.cuh"

namespace nb200 {

constexpr int FRONT_THREADS = 128;
constexpr int FRONT_PTS = 2;     // proposals per thread (register blocking)
constexpr int FRONT_TILE = FRONT_THREADS * FRONT_PTS;

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
  // gather mode (nb200_materialize): proposal i is the GLOBAL proposal index
  // gather[i]; only the row is written (bit-identical to the cycle's)
  const unsigned long long* gather;
  int xs_f16;             // emulator input rows in fp16 (NB200_MLP_F16)
};

__device__ __forceinline__ unsigned long long front_index(const FrontArgs& A,
                                                          long long i) {
  if (A.gather) return __ldg(A.gather + (i < A.n ? i : A.n - 1));
  return A.offset + (unsigned long long)i;
}

// kept out of line so that the four likelihood bodies do not take part in
// the register allocation of the matrix-vector loops
__device__ __noinline__ double front_loglike(int like_id, const double* p,
                                             const double* x, int d) {
  return loglike_eval(like_id, p, x, d);
}

// The kernel is bound by shared-memory wavefronts (ncu: 81 % of the LSU data
// pipe with one proposal per thread), so every thread carries TWO proposals
// through the matrix-vector products: the eight factor entries of a step
// (4 x LDS.128) are loaded once and feed 16 DFMA.
//
// acc_p[r] = sum_{j < jmax} MT[j][i0 + r] * v_pj,
//   v_pj = x_p[j] * s_p          (SCALE: the radial factor of the ball draw)
//   v_pj = x_p[j] - c[j]         (SUBTRACT: whitening)
template <bool SUBTRACT, bool SCALE>
__device__ __forceinline__ void mv8x2(const double* __restrict__ MT, int d8,
                                      int i0, int jmax, const double* x0,
                                      const double* x1,
                                      const double* __restrict__ c, double s0,
                                      double s1, double (&a0)[8],
                                      double (&a1)[8]) {
#pragma unroll
  for (int r = 0; r < 8; ++r) { a0[r] = 0.0; a1[r] = 0.0; }
  const double2* col = reinterpret_cast<const double2*>(MT + i0);
  const int ld = d8 >> 1;   // double2 per transposed row
#pragma unroll 4
  for (int j = 0; j < jmax; ++j) {
    double v0 = x0[j], v1 = x1[j];
    if (SUBTRACT) { const double cj = c[j]; v0 -= cj; v1 -= cj; }
    if (SCALE) { v0 *= s0; v1 *= s1; }
    const double2 m0 = col[j * ld], m1 = col[j * ld + 1];
    const double2 m2 = col[j * ld + 2], m3 = col[j * ld + 3];
    a0[0] = fma(m0.x, v0, a0[0]); a1[0] = fma(m0.x, v1, a1[0]);
    a0[1] = fma(m0.y, v0, a0[1]); a1[1] = fma(m0.y, v1, a1[1]);
    a0[2] = fma(m1.x, v0, a0[2]); a1[2] = fma(m1.x, v1, a1[2]);
    a0[3] = fma(m1.y, v0, a0[3]); a1[3] = fma(m1.y, v1, a1[3]);
    a0[4] = fma(m2.x, v0, a0[4]); a1[4] = fma(m2.x, v1, a1[4]);
    a0[5] = fma(m2.y, v0, a0[5]); a1[5] = fma(m2.y, v1, a1[5]);
    a0[6] = fma(m3.x, v0, a0[6]); a1[6] = fma(m3.x, v1, a1[6]);
    a0[7] = fma(m3.y, v0, a0[7]); a1[7] = fma(m3.y, v1, a1[7]);
  }
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
  double* rows = iscaleN + d8;             // FRONT_TILE x stride
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
  // proposal p of this thread is local row p * FRONT_THREADS + threadIdx.x
  // (consecutive threads -> consecutive rows: conflict-free 64-bit accesses)
  double* xr[FRONT_PTS];
#pragma unroll
  for (int p = 0; p < FRONT_PTS; ++p)
    xr[p] = rows + (p * FRONT_THREADS + (int)threadIdx.x) * stride;
  const long long n_tiles = (A.n + FRONT_TILE - 1) / FRONT_TILE;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long base = tile * FRONT_TILE;
    const int nrows = (int)min((long long)FRONT_TILE, A.n - base);
    long long gi[FRONT_PTS];
    bool valid[FRONT_PTS];
    int kq[FRONT_PTS];
    double rq[FRONT_PTS], sq[FRONT_PTS];
    // ---- normals -> direction and radial factor of a uniform point of the
    // unit ball (basic.py:376-379); the row holds the raw normals, the factor
    // scale = u^(1/d) / |z| is applied on the fly in the product below -------
    // (the two proposals run through the generator side by side, in
    // straight-line code, so that their dependent integer chains interleave;
    // rows past the end of the batch are computed and dropped)
    {
      double n2[FRONT_PTS];
      double uq[FRONT_PTS];
      const Philox rng0(front_index(A, base + threadIdx.x), A.stream_id,
                        A.seed);
      const Philox rng1(front_index(A, base + FRONT_THREADS + threadIdx.x),
                        A.stream_id, A.seed);
#pragma unroll
      for (int p = 0; p < FRONT_PTS; ++p) {
        const int li = p * FRONT_THREADS + (int)threadIdx.x;
        gi[p] = base + li;
        valid[p] = li < nrows;
        const uint4 w0 = p == 0 ? rng0.block(0) : rng1.block(0);
        const double uk = u01_32(w0.x);
        int k = 0;
        while (k < K - 1 && !(uk < __ldg(cdf + k))) ++k;
        kq[p] = k;
        rq[p] = u01_32(w0.y);
        uq[p] = u01_53(w0.z, w0.w);
        n2[p] = 0.0;
      }
      for (int j = 0; j < d; j += 4) {
        const uint4 wa = rng0.block(1 + (j >> 2));
        const uint4 wb = rng1.block(1 + (j >> 2));
        float ga[4], gb[4];
        normal2(wa.x, wa.y, ga[0], ga[1]);
        normal2(wb.x, wb.y, gb[0], gb[1]);
        normal2(wa.z, wa.w, ga[2], ga[3]);
        normal2(wb.z, wb.w, gb[2], gb[3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (j + q < d) {
            const double va = (double)ga[q], vb = (double)gb[q];
            xr[0][j + q] = va;
            xr[1][j + q] = vb;
            n2[0] = fma(va, va, n2[0]);
            n2[1] = fma(vb, vb, n2[1]);
          }
        }
      }
#pragma unroll
      for (int p = 0; p < FRONT_PTS; ++p)
        sq[p] = pow(uq[p], 1.0 / (double)d) / sqrt(n2[p]);
    }
    // ---- x = B (z * scale) + c in place (basic.py:380 -> :342), highest row
    // group first; the unit-cube test (union.py:313-314) runs on the
    // registers as the coordinates are produced ------------------------------
    bool in_cube[FRONT_PTS] = {true, true};
    if (kq[0] == kq[1]) {
      const double* Bk = BT + (size_t)kq[0] * mat;
      const double* ck = cK + kq[0] * d8;
      for (int i0 = d8 - 8; i0 >= 0; i0 -= 8) {
        double a0[8], a1[8];
        mv8x2<false, true>(Bk, d8, i0, min(i0 + 8, d), xr[0], xr[1], nullptr,
                           sq[0], sq[1], a0, a1);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (i0 + q < d) {
            const double cq = ck[i0 + q];
            const double v0 = a0[q] + cq, v1 = a1[q] + cq;
            xr[0][i0 + q] = v0; xr[1][i0 + q] = v1;
            in_cube[0] = in_cube[0] && (v0 >= 0.0) && (v0 < 1.0);
            in_cube[1] = in_cube[1] && (v1 >= 0.0) && (v1 < 1.0);
          }
        }
      }
    } else {
      // the two proposals picked different ellipsoids: one factor each (the
      // second operand of the pair is a dummy)
#pragma unroll
      for (int p = 0; p < FRONT_PTS; ++p) {
        const double* Bk = BT + (size_t)kq[p] * mat;
        const double* ck = cK + kq[p] * d8;
        for (int i0 = d8 - 8; i0 >= 0; i0 -= 8) {
          double a0[8], a1[8];
          mv8x2<false, true>(Bk, d8, i0, min(i0 + 8, d), xr[p], xr[p], nullptr,
                             sq[p], sq[p], a0, a1);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (i0 + q < d) {
              const double v0 = a0[q] + ck[i0 + q];
              xr[p][i0 + q] = v0;
              in_cube[p] = in_cube[p] && (v0 >= 0.0) && (v0 < 1.0);
            }
          }
        }
      }
    }
    // ---- dispositions ------------------------------------------------------
    uint8_t cd[FRONT_PTS];
    bool alive[FRONT_PTS];    // still needs ellipsoid tests
#pragma unroll
    for (int p = 0; p < FRONT_PTS; ++p) {
      cd[p] = NB200_CODE_IN_SHELL;
      if (A.unit && !in_cube[p]) cd[p] = NB200_CODE_CUBE_REJECT;
      alive[p] = valid[p] && cd[p] == NB200_CODE_IN_SHELL && !A.gather;
    }
    // whitening w.r.t. the neural bound's ellipsoid: r2 and the standardised,
    // tf32-rounded emulator input row (with the constant-one bias column at
    // index d), for the proposals flagged in `want`
    double r2_nb[FRONT_PTS] = {-1.0, -1.0};
    auto whiten_store = [&](const bool (&want)[FRONT_PTS]) {
      if (!(want[0] || want[1])) return;
      // row pointers in 32-bit words: k0p tf32 values or k0p / 2 fp16 pairs
      const int xw = A.xs_f16 ? A.k0p >> 1 : A.k0p;
      uint32_t* xrow0 = reinterpret_cast<uint32_t*>(xs32) + gi[0] * (long long)xw;
      uint32_t* xrow1 = reinterpret_cast<uint32_t*>(xs32) + gi[1] * (long long)xw;
      double r20 = 0.0, r21 = 0.0;
      for (int i0 = 0; i0 < A.k0p; i0 += 8) {
        float f0[8], f1[8];
        if (i0 < d8) {
          double a0[8], a1[8];
          mv8x2<true, false>(nbT, d8, i0, A.lower_n ? min(i0 + 8, d) : d,
                             xr[0], xr[1], cN, 0.0, 0.0, a0, a1);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            r20 = fma(a0[q], a0[q], r20);
            r21 = fma(a1[q], a1[q], r21);
            const double mq = meanN[i0 + q], iq = iscaleN[i0 + q];
            f0[q] = (float)((a0[q] - mq) * iq);
            f1[q] = (float)((a1[q] - mq) * iq);
            if (i0 + q == d) { f0[q] = 1.0f; f1[q] = 1.0f; }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            f0[q] = f1[q] = (i0 + q == d) ? 1.0f : 0.0f;
        }
        if (A.xs_f16) {
          if (want[0])
            *reinterpret_cast<uint4*>(xrow0 + (i0 >> 1)) = make_uint4(
                pack_f16x2(f0[0], f0[1]), pack_f16x2(f0[2], f0[3]),
                pack_f16x2(f0[4], f0[5]), pack_f16x2(f0[6], f0[7]));
          if (want[1])
            *reinterpret_cast<uint4*>(xrow1 + (i0 >> 1)) = make_uint4(
                pack_f16x2(f1[0], f1[1]), pack_f16x2(f1[2], f1[3]),
                pack_f16x2(f1[4], f1[5]), pack_f16x2(f1[6], f1[7]));
          continue;
        }
        uint32_t pk0[8], pk1[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(pk0[q]) : "f"(f0[q]));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(pk1[q]) : "f"(f1[q]));
        }
        if (want[0]) {
          *reinterpret_cast<uint4*>(xrow0 + i0) =
              make_uint4(pk0[0], pk0[1], pk0[2], pk0[3]);
          *reinterpret_cast<uint4*>(xrow0 + i0 + 4) =
              make_uint4(pk0[4], pk0[5], pk0[6], pk0[7]);
        }
        if (want[1]) {
          *reinterpret_cast<uint4*>(xrow1 + i0) =
              make_uint4(pk1[0], pk1[1], pk1[2], pk1[3]);
          *reinterpret_cast<uint4*>(xrow1 + i0 + 4) =
              make_uint4(pk1[4], pk1[5], pk1[6], pk1[7]);
        }
      }
      if (want[0]) r2_nb[0] = r20;
      if (want[1]) r2_nb[1] = r21;
    };
    // overlap count n_bound = sum_k contains_k (union.py:316-317)
    if (alive[0] || alive[1]) {
      int nbnd[FRONT_PTS] = {0, 0};
      for (int kk = 0; kk < K; ++kk) {
        double r20, r21;
        if (kk == A.same_k) {
          whiten_store(alive);      // same ellipsoid: one pass serves both
          r20 = r2_nb[0]; r21 = r2_nb[1];
        } else {
          const double* MT = BinvT + (size_t)kk * mat;
          const int lower = (A.lower_k >> kk) & 1;
          r20 = 0.0; r21 = 0.0;
          for (int i0 = 0; i0 < d8; i0 += 8) {
            double a0[8], a1[8];
            mv8x2<true, false>(MT, d8, i0, lower ? min(i0 + 8, d) : d, xr[0],
                               xr[1], cK + kk * d8, 0.0, 0.0, a0, a1);
#pragma unroll
            for (int q = 0; q < 8; ++q) {      // pad rows contribute 0
              r20 = fma(a0[q], a0[q], r20);
              r21 = fma(a1[q], a1[q], r21);
            }
          }
        }
        nbnd[0] += r20 < 1.0 ? 1 : 0;
        nbnd[1] += r21 < 1.0 ? 1 : 0;
      }
#pragma unroll
      for (int p = 0; p < FRONT_PTS; ++p)
        if (alive[p] && !(rq[p] > 1.0 - 1.0 / (double)nbnd[p])) {
          cd[p] = NB200_CODE_OVERLAP_REJECT;      // union.py:318-319
          alive[p] = false;
        }
      // NeuralBound: ellipsoid test (neural.py:117)
      if (A.same_k < 0) whiten_store(alive);
    }
#pragma unroll
    for (int p = 0; p < FRONT_PTS; ++p) {
      if (!valid[p] || A.gather) continue;
      bool in_ell = false;
      if (alive[p]) {
        in_ell = r2_nb[p] < 1.0;
        if (!in_ell) cd[p] = NB200_CODE_NN_REJECT;
      }
      code[gi[p]] = cd[p];
      maskj[gi[p]] = in_ell ? 1 : 0;
      if (A.log_l)
        A.log_l[gi[p]] = cd[p] == NB200_CODE_IN_SHELL
                             ? front_loglike(A.like_id, A.like_p, xr[p], d)
                             : nan("");
    }
    __syncthreads();
    store_rows(points, base, nrows, d, stride, rows);
    __syncthreads();
  }
}

// Host side: does the fast path apply, and launch.
struct FrontMmaArgs;
bool front_mma_applicable(const int32_t* meta_h, int bound, size_t* smem_out,
                          FrontMmaArgs* args);
int launch_front_mma(const int32_t* meta_h, const int32_t* meta_d,
                     const double* data_d, int bound, int64_t n, uint64_t seed,
                     uint64_t offset, uint32_t stream_id, double* points,
                     uint8_t* code, uint8_t* maskj, float* xs32, int like_id,
                     const double* like_p, double* log_l,
                     const unsigned long long* gather, int xs_f16,
                     cudaStream_t st);

// NB200_FRONT=dfma forces the DFMA kernel (A/B measurements, tests)
static bool front_mma_wanted() {
  const char* e = getenv("NB200_FRONT");
  return !(e && strcmp(e, "dfma") == 0);
}

static bool front2_applicable(const int32_t* meta_h, int bound,
                              size_t* smem_out, FrontArgs* args) {
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
                         (size_t)FRONT_TILE * (d | 1);
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

// one ellipsoid: the DMMA kernel; several: the DFMA kernel above
bool front_applicable(const int32_t* meta_h, int bound, size_t* smem_out,
                      FrontArgs* args) {
  if (front_mma_wanted() &&
      front_mma_applicable(meta_h, bound, nullptr, nullptr))
    return true;
  return front2_applicable(meta_h, bound, smem_out, args);
}

int launch_front(const int32_t* meta_h, const int32_t* meta_d,
                 const double* data_d, int bound, int64_t n, uint64_t seed,
                 uint64_t offset, uint32_t stream_id, double* points,
                 uint8_t* code, uint8_t* maskj, float* xs32, int like_id,
                 const double* like_p, double* log_l,
                 const unsigned long long* gather, int xs_f16,
                 cudaStream_t st) {
  if (front_mma_wanted() &&
      front_mma_applicable(meta_h, bound, nullptr, nullptr))
    return launch_front_mma(meta_h, meta_d, data_d, bound, n, seed, offset,
                            stream_id, points, code, maskj, xs32, like_id,
                            like_p, log_l, gather, xs_f16, st);
  FrontArgs A;
  size_t smem = 0;
  NB_CHECK(front2_applicable(meta_h, bound, &smem, &A), "front kernel n/a");
  A.seed = seed; A.offset = offset; A.stream_id = stream_id; A.n = n;
  A.like_id = like_id; A.like_p = like_p; A.log_l = log_l;
  A.gather = gather;
  A.xs_f16 = xs_f16;
  if (xs_f16) A.k0p = (A.d + 1 + 15) / 16 * 16;
  NB_CUDA(cudaFuncSetAttribute(k_front,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = (n + FRONT_TILE - 1) / FRONT_TILE;
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
