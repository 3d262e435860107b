// Fused fp64 front end of the cycle on the FP64 tensor cores (DMMA), for the
// common bound with ONE ellipsoid (K == 1, a plain Ellipsoid mixture) and one
// neural bound: the same job as k_front in nb200_front.cu --
//   Union.sample draw (nautilus/bounds/union.py:305-319, basic.py:376-381),
//   unit-cube cut, overlap acceptance, NeuralBound's ellipsoid test and
//   whitening (bounds/neural.py:117-119), emulator input standardisation
//   (nautilus/neural.py:115), optional built-in likelihood --
// but the two matrix-vector products per proposal, x = B z + c and
// t = B_inv (x - c), run as mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4): measured
// here at 37.0 TFLOP/s, the same rate as the DFMA pipe, for 1/8 of the issue
// slots and no shared-memory traffic per FMA (k_front was bound by LDS
// wavefronts, then by issue latency).
//
// Mapping: a warp owns 32 proposals whose rows live in its private slice of
// shared memory (row stride S = d8 + 4 doubles: conflict-free A-fragment
// loads); lane = (p, q) = (lane / 4, lane % 4).  For every group g of 8
// proposals
//   A fragment  a[p][q]      = row (8g + p), column 4 Kb + q
//   B fragment  b[q][n=p]    = M[8 I + p][4 Kb + q]   (staged in this order)
//   C fragment  c[p][2q..]   = output rows 8 I + 2q, 8 I + 2q + 1 of row 8g+p
// Lower-triangular factors skip the blocks above the diagonal (exact zeros).
// The generator is spread the same way: lane (p, q) draws the Philox blocks
// b = q, q + 4, ... (4 normals each) of the proposals 8g + p, and owns
// proposal 8q + p for everything scalar (acceptance uniform, radial factor,
// disposition, likelihood).  The warps of a CTA never synchronise with each
// other inside the loop.
#include <math.h>
#include <stdlib.h>

#include "nb200_device.cuh"
#include "nb200_rng.cuh"
#include "nb200_tc.cuh"

namespace nb200 {

// warps per CTA: 8 for rows up to 64 wide (two CTAs per SM); wider rows
// (d up to 128, e.g. BASELINE config 5) keep fewer rows on chip
__host__ __device__ constexpr int fm_warps(int d8) {
  return d8 <= 64 ? 8 : d8 <= 112 ? 4 : 2;
}
// Factors wider than 64 are staged TRIANGULAR-PACKED (they must be exactly
// lower-triangular): row block I keeps its 2 I + 2 k-blocks only, so both
// factors of a 100-D ellipsoid take 93 KB instead of 173 KB.
__host__ __device__ constexpr bool fm_tri(int d8) { return d8 > 64; }
__host__ __device__ constexpr int fm_fac_doubles(int d8) {
  return fm_tri(d8) ? (d8 / 8) * (d8 / 8 + 1) * 32 : d8 * d8;
}
// FAST kernels of wide rows stage ONE factor (B): the exact whitening runs
// for one tile in a million and fetches B_inv from global memory, so the
// shared memory of the second factor holds more rows instead -- 6 warps
// instead of 4 at d = 100.  Shared memory of such a CTA in doubles, and the
// largest warp count that fits 220 KB:
__host__ __device__ constexpr size_t fm_fast_wide_doubles(int d8, int warps) {
  return (size_t)fm_fac_doubles(d8) + 6 * (size_t)d8 +
         (size_t)warps * 32 * (d8 + 4);
}
__host__ __device__ constexpr int fm_warps(int d8, bool fast) {
  if (!(fast && fm_tri(d8))) return fm_warps(d8);
  for (int w = 8; w > 2; w -= 2)
    if (fm_fast_wide_doubles(d8, w) * 8 <= 220 * 1024) return w;
  return 2;
}
// first fragment (in units of 32 doubles) of row block I
template <int D8>
__device__ __forceinline__ int fm_row_block(int I) {
  return fm_tri(D8) ? I * (I + 1) : I * (D8 / 4);
}

struct FrontMmaArgs {
  int rec_off, d, d8, unit, k0p, S;
  int lower_b, lower_n;   // B_inv of the mixture / of the neural bound is
                          // exactly lower-triangular
  int same;               // neural bound's ellipsoid == the mixture's
  unsigned long long seed, offset;
  unsigned int stream_id;
  long long n;
  int like_id;
  const double* like_p;
  double* log_l;
  // gather mode (nb200_materialize): proposal i of the launch is the GLOBAL
  // proposal index gather[i] instead of offset + i, and only the row is
  // written -- same instructions in the same order as the cycle's launch, so
  // the regenerated row is bit-identical to the one the cycle produced
  const unsigned long long* gather;
  // emulator input rows in fp16 (NB200_MLP_F16): k0p halves per row, two to
  // a 32-bit word, instead of k0p tf32 words
  int xs_f16;
  // whitening shortcut (FAST kernels; the neural bound's ellipsoid is the
  // mixture's): half-width of the guard band around r^2 = 1 inside which the
  // exact whitening is redone (see k_front_mma)
  double fast_tau;
};

// global proposal index (the Philox counter) of local proposal i
__device__ __forceinline__ unsigned long long fm_index(const FrontMmaArgs& A,
                                                       long long i) {
  if (A.gather) return __ldg(A.gather + (i < A.n ? i : A.n - 1));
  return A.offset + (unsigned long long)i;
}

__device__ __noinline__ double front_mma_loglike(int like_id, const double* p,
                                                 const double* x, int d) {
  return loglike_eval(like_id, p, x, d);
}

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 "
      "{%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// T[g] = M (rows 8I..8I+7) . (row_g - c) for the four proposal groups; all
// trip counts and strides are compile-time, so the loop unrolls into
// LDS / DMMA with immediate offsets (the first version spent more issue
// slots on address arithmetic than on DMMAs)
// GLOBAL_B: the B fragments come straight from the row-major matrix in global
// memory (frag_lane = M, this lane's element of fragment (I, Kb) is
// M[8 I + lane / 4][4 Kb + lane % 4]) -- the rarely taken exact whitening of
// the wide FAST kernels; same values, same DMMA sequence as the staged form.
template <int D8, bool SUBTRACT, bool GLOBAL_B = false>
__device__ __forceinline__ void mma_rows(const double* __restrict__ frag_lane,
                                         int I, int kb_end,
                                         const double* row_lane,
                                         const double* __restrict__ c_lane,
                                         double (&T)[4][2], int d = 0) {
  constexpr int NK = D8 / 4, S = D8 + 4;
  const int gl = threadIdx.x & 31;
  const int gi_row = 8 * I + (gl >> 2);
  if (!GLOBAL_B) frag_lane += fm_row_block<D8>(I) * 32;
#pragma unroll
  for (int g = 0; g < 4; ++g) { T[g][0] = 0.0; T[g][1] = 0.0; }
#pragma unroll
  for (int Kb = 0; Kb < NK; ++Kb) {
    if (Kb < kb_end) {
      double b;
      if (GLOBAL_B) {
        const int j = 4 * Kb + (gl & 3);
        b = (gi_row < d && j < d) ? __ldg(frag_lane + (size_t)gi_row * d + j)
                                  : 0.0;
      } else {
        b = frag_lane[Kb * 32];
      }
      const double cq = SUBTRACT ? c_lane[4 * Kb] : 0.0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        double a = row_lane[g * 8 * S + 4 * Kb];
        if (SUBTRACT) a -= cq;
        dmma(T[g], a, b);
      }
    }
  }
}

// Columns i0, i0 + 1 of the emulator's input row of proposal gi from the
// whitened coordinates t0, t1: standardised, rounded, the constant-one bias
// column at index d (pack_tc), zeros behind it.
__device__ __forceinline__ void emit_xs(const FrontMmaArgs& A, float* xs32,
                                        long long gi, int i0, int d, double t0,
                                        double t1, double m0, double m1,
                                        double s0, double s1, bool wanted) {
  float v0 = (float)((t0 - m0) * s0);
  float v1 = (float)((t1 - m1) * s1);
  if (i0 >= d) v0 = i0 == d ? 1.0f : 0.0f;
  if (i0 + 1 >= d) v1 = i0 + 1 == d ? 1.0f : 0.0f;
  if (A.xs_f16) {
    if (wanted)
      reinterpret_cast<uint32_t*>(xs32)[gi * (long long)(A.k0p >> 1) +
                                        (i0 >> 1)] = pack_f16x2(v0, v1);
  } else {
    uint32_t k0, k1;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(k0) : "f"(v0));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(k1) : "f"(v1));
    if (wanted)
      *reinterpret_cast<uint2*>(xs32 + gi * (long long)A.k0p + i0) =
          make_uint2(k0, k1);
  }
}

// The same from z and the radial factor (FAST kernels): (s z - mean) / scale
// as one multiply and one FMA, the padding selects only where the row ends.
__device__ __forceinline__ void emit_xs_fast(const FrontMmaArgs& A,
                                             float* xs32, long long gi, int i0,
                                             int d, bool tail, double t0,
                                             double t1, double s0, double s1,
                                             double nms0, double nms1,
                                             bool wanted) {
  float v0 = (float)fma(t0, s0, nms0);
  float v1 = (float)fma(t1, s1, nms1);
  if (tail) {
    if (i0 >= d) v0 = i0 == d ? 1.0f : 0.0f;
    if (i0 + 1 >= d) v1 = i0 + 1 == d ? 1.0f : 0.0f;
  }
  if (!wanted) return;
  if (A.xs_f16) {
    reinterpret_cast<uint32_t*>(xs32)[gi * (long long)(A.k0p >> 1) +
                                      (i0 >> 1)] = pack_f16x2(v0, v1);
  } else {
    uint32_t k0, k1;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(k0) : "f"(v0));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(k1) : "f"(v1));
    *reinterpret_cast<uint2*>(xs32 + gi * (long long)A.k0p + i0) =
        make_uint2(k0, k1);
  }
}

// FAST (the neural bound's ellipsoid IS the mixture's, the usual unimodal
// bound): the whitened point t = B_inv (x - c) of x = c + s B z is s z up to
// the rounding of the round trip, and its squared radius is s^2 |z|^2.  The
// emulator's input row -- rounded to fp16 / tf32 anyway -- is taken from s z
// while z is still in shared memory, and the second pass of DMMAs (half of
// the kernel's tensor work) is skipped UNLESS some proposal of the tile has
// s^2 |z|^2 within fast_tau of 1: then the exact whitening decides, for the
// whole tile, as in the other kernels.  fast_tau is 256 d eps times the
// rounding amplification of the round trip (amp_log2 in the record header),
// at least 1e-9, so the membership decisions stay those of the exact
// arithmetic; a tile takes the exact path with probability ~16 d fast_tau.
template <int D8, bool FAST>
__global__ void __launch_bounds__(fm_warps(D8, FAST) * 32, D8 <= 64 ? 2 : 1)
k_front_mma(const FrontMmaArgs A, const int32_t* __restrict__ meta,
            const double* __restrict__ data, double* __restrict__ points,
            uint8_t* __restrict__ code, uint8_t* __restrict__ maskj,
            float* __restrict__ xs32) {
  extern __shared__ __align__(16) double sm[];
  const Rec rec{meta + A.rec_off};
  constexpr int d8 = D8, S = D8 + 4, nI = D8 / 8, nK = D8 / 4;
  constexpr int FM_WARPS = fm_warps(D8, FAST), FM_THREADS = FM_WARPS * 32;
  constexpr bool TRI = fm_tri(D8);
  // wide FAST kernels keep B only; B_inv is read from global memory by the
  // (rare) exact whitening
  constexpr bool ONE_FAC = FAST && TRI;
  const int d = A.d;
  constexpr int fsz = fm_fac_doubles(D8);  // doubles per staged factor
  double* fB = sm;                         // B        in fragment order
  double* fBinv = fB + (ONE_FAC ? 0 : fsz); // B_inv (mixture)
  double* fN = fBinv + fsz;                // B_inv (neural), if different
  double* cM = fN + (A.same ? 0 : fsz);    // d8: centre of the mixture
  double* cN = cM + d8;                    // d8: centre of the neural bound
  double* meanN = cN + d8;                 // d8
  double* iscaleN = meanN + d8;            // d8: 1 / scale
  double* nmsN = iscaleN + d8;             // d8: -mean / scale
  double* muL = nmsN + d8;                 // d8: centre of a Gaussian likelihood
  double* rows_all = muL + d8;             // FM_WARPS x 32 x S
  const int32_t* nb = rec.nb(0);
  const int32_t* mix = rec.mix(0);

  // factors -> fragment order: frag[I][Kb][lane] = M[8I + lane/4][4Kb + lane%4]
  const int n_fac = ONE_FAC ? 1 : A.same ? 2 : 3;
  // (four independent loads in flight per thread: the staging is a chain of
  // L2 round trips otherwise, ~5 % of the kernel at 2^20 proposals)
  for (int e0 = threadIdx.x; e0 < n_fac * d8 * d8; e0 += 4 * FM_THREADS) {
    double val[4];
    int dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * FM_THREADS;
      dst[u] = -1;
      val[u] = 0.0;
      if (e >= n_fac * d8 * d8) continue;
      const int which = e / (d8 * d8), rem = e - which * (d8 * d8);
      const int blk = rem >> 5, ln = rem & 31;
      const int I = blk / nK, Kb = blk - I * nK;
      if (TRI && Kb >= 2 * I + 2) continue;  // above the diagonal: not kept
      const int i = 8 * I + (ln >> 2), j = 4 * Kb + (ln & 3);
      const double* src = data + (which == 0 ? mix[4] : which == 1 ? mix[5]
                                                                   : nb[1]);
      dst[u] = which * fsz + (fm_row_block<D8>(I) + Kb) * 32 + ln;
      if (i < d && j < d) val[u] = __ldg(src + (size_t)i * d + j);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (dst[u] >= 0) sm[dst[u]] = val[u];
  }
  // a Gaussian likelihood is summed in the fragment layout, next to the
  // unit-cube test (four lanes share a proposal's row); the others walk the
  // finished row
  const bool gauss = A.log_l != nullptr && A.like_id == NB200_LIKE_GAUSSIAN;
  for (int e = threadIdx.x; e < 6 * d8; e += FM_THREADS) {
    const int which = e / d8, i = e - which * d8;
    double v = 0.0;
    if (i < d) {
      if (which == 0) v = data[mix[3] + i];
      else if (which == 1) v = data[nb[0] + i];
      else if (which == 2) v = data[nb[5] + i];
      else if (which == 3) v = 1.0 / data[nb[6] + i];
      else if (which == 4) v = -data[nb[5] + i] * (1.0 / data[nb[6] + i]);
      else if (gauss) v = A.like_p[2 + i];
    }
    cM[e] = v;
  }
  __syncthreads();
  const double* fNeural = A.same ? fBinv : fN;
  const int lower_n = A.same ? A.lower_b : A.lower_n;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = lane >> 2, q = lane & 3;
  double* rows = rows_all + (size_t)warp * 32 * S;
  const int nblk = (d + 3) >> 2;           // Philox blocks of 4 normals
  const double inv_d = 1.0 / (double)d;
  const long long n_tiles = (A.n + 31) / 32;
  const long long tile_step = (long long)gridDim.x * FM_WARPS;

  for (long long tile = (long long)blockIdx.x * FM_WARPS + warp;
       tile < n_tiles; tile += tile_step) {
    const long long base = tile * 32;
    // ---- generator ---------------------------------------------------------
    // this lane's own proposal: 8q + p
    const long long gi_own = base + 8 * q + p;
    const Philox rng_own(fm_index(A, gi_own), A.stream_id, A.seed);
    const uint4 w0 = rng_own.block(0);
    const double r_own = u01_32(w0.y);     // acceptance uniform (K == 1: the
    const double u_own = u01_53(w0.z, w0.w);   // ellipsoid choice is moot)
    double n2g[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const Philox rng(fm_index(A, base + 8 * g + p), A.stream_id, A.seed);
      double acc = 0.0;
      double* zr = rows + (8 * g + p) * S;
      // nK blocks cover the padded row; compile-time trip count, so that the
      // Philox chains of all blocks and groups interleave (ILP)
#pragma unroll
      for (int bb = 0; bb < (nK + 3) / 4; ++bb) {
        const int b = q + 4 * bb;
        if (b >= nK) continue;
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (b < nblk) {
          const uint4 w = rng.block(1 + b);
          float g0, g1, g2, g3;
          normal2(w.x, w.y, g0, g1);
          normal2(w.z, w.w, g2, g3);
          v[0] = (double)g0; v[1] = (double)g1;
          v[2] = (double)g2; v[3] = (double)g3;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (4 * b + t >= d) v[t] = 0.0;
            acc = fma(v[t], v[t], acc);
          }
        }
        *reinterpret_cast<double2*>(zr + 4 * b) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2*>(zr + 4 * b + 2) = make_double2(v[2], v[3]);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      n2g[g] = acc;
    }
    const double n2_own = q == 0 ? n2g[0] : q == 1 ? n2g[1]
                                                   : q == 2 ? n2g[2] : n2g[3];
    // radial factor of the uniform ball draw (basic.py:377-379)
    // (u^(1/d) as exp(log(u) / d): ~2e-16 relative, half of pow()'s
    // instructions; u = 0 gives 0 either way)
    const double s_own = exp(log(u_own) * inv_d) * rsqrt(n2_own);
    double sg[4];
#pragma unroll
    for (int g = 0; g < 4; ++g)
      sg[g] = __shfl_sync(0xffffffffu, s_own, 4 * p + g);
    __syncwarp();
    const bool emit = FAST && !A.gather;
    const double r2_fast = s_own * s_own * n2_own;

    // ---- x = s (B z) + c in place, highest row block first (B is lower
    // triangular: block I needs z[k < 8I + 8], which the blocks below have
    // not overwritten); unit-cube test on the fragments ----------------------
    // The coordinates also leave for global memory straight from the
    // fragments: the four lanes of a proposal cover 64 contiguous bytes.
    bool cube[4] = {true, true, true, true};
    double ll2[4] = {0.0, 0.0, 0.0, 0.0};  // sum (x - mu)^2, this lane's columns
    const double* row_lane = rows + p * S + q;
    const bool pair_ok = (d & 1) == 0;     // 16-byte stores need even offsets
    // (row-block loops are unrolled only for narrow rows: 13 blocks x 26
    // k-blocks of a 100-D row would not fit the instruction cache)
#pragma unroll(D8 <= 64 ? 8 : 1)
    for (int I = nI - 1; I >= 0; --I) {
      double T[4][2];
      mma_rows<D8, false>(fB + lane, I, 2 * I + 2, row_lane, nullptr, T);
      const int i0 = 8 * I + 2 * q;
      if (emit) {
        // this lane's columns of z, before x takes their place
        const double2 sc = *reinterpret_cast<const double2*>(iscaleN + i0);
        const double2 nm = *reinterpret_cast<const double2*>(nmsN + i0);
        const bool tail = 8 * I + 8 > d;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const double2 zz = *reinterpret_cast<const double2*>(
              rows + (8 * g + p) * S + i0);
          const long long gi = base + 8 * g + p;
          emit_xs_fast(A, xs32, gi, i0, d, tail, sg[g] * zz.x, sg[g] * zz.y,
                       sc.x, sc.y, nm.x, nm.y, gi < A.n);
        }
      }
      __syncwarp();              // every lane has read this block's columns
      const double c0 = cM[i0], c1 = cM[i0 + 1];
      const double mu0 = muL[i0], mu1 = muL[i0 + 1];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const double x0 = fma(sg[g], T[g][0], c0);
        const double x1 = fma(sg[g], T[g][1], c1);
        *reinterpret_cast<double2*>(rows + (8 * g + p) * S + i0) =
            make_double2(x0, x1);
        if (i0 < d) cube[g] = cube[g] && (x0 >= 0.0) && (x0 < 1.0);
        if (i0 + 1 < d) cube[g] = cube[g] && (x1 >= 0.0) && (x1 < 1.0);
        if (gauss) {
          // (padding columns: x = 0 and mu = 0)
          const double e0 = x0 - mu0, e1 = x1 - mu1;
          ll2[g] = fma(e1, e1, fma(e0, e0, ll2[g]));
        }
        const long long gi = base + 8 * g + p;
        if (gi < A.n) {
          double* dst = points + gi * (long long)d + i0;
          if (pair_ok && i0 + 1 < d) {
            *reinterpret_cast<double2*>(dst) = make_double2(x0, x1);
          } else {
            if (i0 < d) dst[0] = x0;
            if (i0 + 1 < d) dst[1] = x1;
          }
        }
      }
    }
    unsigned cube_bits = 0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      unsigned ok = cube[g] ? 1u : 0u;
      ok &= __shfl_xor_sync(0xffffffffu, ok, 1);
      ok &= __shfl_xor_sync(0xffffffffu, ok, 2);
      cube_bits |= ok << g;
    }
    double ll2_own = 0.0;
    if (gauss) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        ll2[g] += __shfl_xor_sync(0xffffffffu, ll2[g], 1);
        ll2[g] += __shfl_xor_sync(0xffffffffu, ll2[g], 2);
      }
      ll2_own = q == 0 ? ll2[0] : q == 1 ? ll2[1] : q == 2 ? ll2[2] : ll2[3];
    }
    __syncwarp();
    if (A.gather) continue;      // materialize: the row is all that is wanted

    double r2m_own, r2n_own;
    bool exact = true;
    if (FAST) {
      // input columns behind the padded row width (bias / zeros)
      for (int I = nI; 8 * I < A.k0p; ++I) {
        const int i0 = 8 * I + 2 * q;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const long long gi = base + 8 * g + p;
          emit_xs(A, xs32, gi, i0, d, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
                  gi < A.n);
        }
      }
      exact = __any_sync(0xffffffffu, fabs(r2_fast - 1.0) < A.fast_tau);
      r2m_own = r2n_own = r2_fast;
    }
    if (exact) {
    // ---- whitening(s): squared radii, emulator input rows -------------------
    double r2m[4] = {0.0, 0.0, 0.0, 0.0};   // w.r.t. the mixture's ellipsoid
    double r2n[4] = {0.0, 0.0, 0.0, 0.0};   // w.r.t. the neural bound's
    if (!A.same) {
#pragma unroll(D8 <= 64 ? 8 : 1)
      for (int I = 0; I < nI; ++I) {
        double T[4][2];
        mma_rows<D8, true>(fBinv + lane, I, A.lower_b ? 2 * I + 2 : nK,
                           row_lane, cM + q, T);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          r2m[g] = fma(T[g][1], T[g][1], fma(T[g][0], T[g][0], r2m[g]));
      }
    }
#pragma unroll(D8 <= 64 ? 10 : 1)
    for (int I = 0; I < nI + 2; ++I) {
      if (8 * I >= A.k0p) break;
      double T[4][2];
      if (I < nI) {
        if (ONE_FAC)
          mma_rows<D8, true, true>(data + nb[1], I, 2 * I + 2, row_lane,
                                   cN + q, T, d);
        else
          mma_rows<D8, true>(fNeural + lane, I, lower_n ? 2 * I + 2 : nK,
                             row_lane, cN + q, T);
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) { T[g][0] = 0.0; T[g][1] = 0.0; }
      }
      const int i0 = 8 * I + 2 * q;
      const double m0 = i0 < d8 ? meanN[i0] : 0.0;
      const double m1 = i0 < d8 ? meanN[i0 + 1] : 0.0;
      const double s0 = i0 < d8 ? iscaleN[i0] : 0.0;
      const double s1 = i0 < d8 ? iscaleN[i0 + 1] : 0.0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        r2n[g] = fma(T[g][1], T[g][1], fma(T[g][0], T[g][0], r2n[g]));
        const long long gi = base + 8 * g + p;
        // rows cut by the unit cube are never looked at by the emulator
        const bool wanted =
            gi < A.n && (!A.unit || ((cube_bits >> g) & 1u));
        emit_xs(A, xs32, gi, i0, d, T[g][0], T[g][1], m0, m1, s0, s1,
                wanted || (FAST && gi < A.n));
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      r2n[g] += __shfl_xor_sync(0xffffffffu, r2n[g], 1);
      r2n[g] += __shfl_xor_sync(0xffffffffu, r2n[g], 2);
      if (!A.same) {
        r2m[g] += __shfl_xor_sync(0xffffffffu, r2m[g], 1);
        r2m[g] += __shfl_xor_sync(0xffffffffu, r2m[g], 2);
      } else {
        r2m[g] = r2n[g];
      }
    }
    r2m_own = q == 0 ? r2m[0] : q == 1 ? r2m[1] : q == 2 ? r2m[2] : r2m[3];
    r2n_own = q == 0 ? r2n[0] : q == 1 ? r2n[1] : q == 2 ? r2n[2] : r2n[3];
    }
    // ---- disposition of this lane's own proposal (8q + p) -------------------
    if (gi_own < A.n) {
      uint8_t cd = NB200_CODE_IN_SHELL;
      bool in_ell = false;
      if (A.unit && !((cube_bits >> q) & 1u)) {
        cd = NB200_CODE_CUBE_REJECT;                   // union.py:313-314
      } else {
        const int nbnd = r2m_own < 1.0 ? 1 : 0;        // union.py:316-317
        if (!(r_own > 1.0 - 1.0 / (double)nbnd)) {     // union.py:318-319
          cd = NB200_CODE_OVERLAP_REJECT;
        } else {
          in_ell = r2n_own < 1.0;                      // neural.py:117
          if (!in_ell) cd = NB200_CODE_NN_REJECT;
        }
      }
      code[gi_own] = cd;
      maskj[gi_own] = in_ell ? 1 : 0;
      if (A.log_l)
        A.log_l[gi_own] =
            cd != NB200_CODE_IN_SHELL ? nan("")
            : gauss ? -0.5 * __ldg(A.like_p) * ll2_own + __ldg(A.like_p + 1)
                    : front_mma_loglike(A.like_id, A.like_p,
                                        rows + (8 * q + p) * S, d);
    }
    __syncwarp();
  }
}

// Host side.
bool front_mma_applicable(const int32_t* meta_h, int bound, size_t* smem_out,
                          FrontMmaArgs* args) {
  const Rec rec = record(meta_h, bound);
  if (rec.kind() != 1 || rec.J() != 1 || rec.K() != 1) return false;
  const int32_t* nb = rec.nb(0);
  if (nb[3] <= 0 || nb[10] < 0 || nb[11] <= 0) return false;
  const int d = rec.d();
  if (rec.mix(0)[1] != 0 || rec.mix(0)[0] != d) return false;
  const int d8 = (d + 7) / 8 * 8;
  if (d8 > 128) return false;
  const int S = d8 + 4;
  const int same = rec.r[10] - 1 == 0;
  // wide rows: triangular-packed factors, which must be exactly lower
  if (fm_tri(d8) && !(rec.mix(0)[6] && (same || nb[2]))) return false;
  const size_t doubles = (size_t)(same ? 2 : 3) * fm_fac_doubles(d8) +
                         6 * (size_t)d8 + (size_t)fm_warps(d8) * 32 * S;
  if (doubles * 8 > (d8 <= 64 ? 200 : 227) * 1024) return false;
  if (smem_out) *smem_out = doubles * 8;
  if (args) {
    args->rec_off = (int)(rec.r - meta_h);
    args->d = d; args->d8 = d8; args->unit = rec.unit();
    args->k0p = (d + 1 + 7) / 8 * 8; args->S = S;
    args->lower_b = rec.mix(0)[6]; args->lower_n = nb[2];
    args->same = same;
  }
  return true;
}

int launch_front_mma(const int32_t* meta_h, const int32_t* meta_d,
                     const double* data_d, int bound, int64_t n, uint64_t seed,
                     uint64_t offset, uint32_t stream_id, double* points,
                     uint8_t* code, uint8_t* maskj, float* xs32, int like_id,
                     const double* like_p, double* log_l,
                     const unsigned long long* gather, int xs_f16,
                     cudaStream_t st) {
  FrontMmaArgs A;
  size_t smem = 0;
  NB_CHECK(front_mma_applicable(meta_h, bound, &smem, &A),
           "DMMA front kernel n/a");
  A.seed = seed; A.offset = offset; A.stream_id = stream_id; A.n = n;
  A.like_id = like_id; A.like_p = like_p; A.log_l = log_l;
  A.gather = gather;
  A.xs_f16 = xs_f16;
  if (xs_f16) A.k0p = (A.d + 1 + 15) / 16 * 16;
  // whitening shortcut: only when one whitening serves both ellipsoids and
  // the guard band that keeps the decisions exact is narrow
  A.fast_tau = 0.0;
  if (A.same) {
    const Rec rec = record(meta_h, bound);
    double tau = ldexp(256.0 * A.d, -53 + rec.r[11]);
    if (tau < 1e-9) tau = 1e-9;
    if (const char* env = getenv("NB200_FRONT_TAU")) tau = atof(env);
    if (tau > 0.0 && (tau <= 1e-4 || getenv("NB200_FRONT_TAU")))
      A.fast_tau = tau;
  }
  const bool fast = A.fast_tau > 0.0;
  void (*kern)(const FrontMmaArgs, const int32_t*, const double*, double*,
               uint8_t*, uint8_t*, float*) = nullptr;
#define NB_FM_CASE(W) \
  case W: kern = fast ? k_front_mma<W, true> : k_front_mma<W, false>; break;
  switch (A.d8) {
    NB_FM_CASE(8) NB_FM_CASE(16) NB_FM_CASE(24) NB_FM_CASE(32)
    NB_FM_CASE(40) NB_FM_CASE(48) NB_FM_CASE(56) NB_FM_CASE(64)
    NB_FM_CASE(72) NB_FM_CASE(80) NB_FM_CASE(88) NB_FM_CASE(96)
    NB_FM_CASE(104) NB_FM_CASE(112) NB_FM_CASE(120) NB_FM_CASE(128)
    default: NB_CHECK(false, "DMMA front kernel: n_dim > 128");
  }
#undef NB_FM_CASE
  const int FM_WARPS = fm_warps(A.d8, fast), FM_THREADS = FM_WARPS * 32;
  if (fast && fm_tri(A.d8))     // one staged factor, more rows on chip
    smem = fm_fast_wide_doubles(A.d8, FM_WARPS) * 8;
  NB_CUDA(cudaFuncSetAttribute(kern,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = (n + 31) / 32;
  const int per_sm = smem * 2 <= 220 * 1024 ? 2 : 1;
  int64_t grid = (int64_t)per_sm * sms;
  const int64_t need = (n_tiles + FM_WARPS - 1) / FM_WARPS;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  ProfScope prof(ST_FUSED, st);
  kern<<<(unsigned)grid, FM_THREADS, smem, st>>>(A, meta_d, data_d, points,
                                                 code, maskj, xs32);
  NB_LAUNCH_OK();
  return 0;
}

}  // namespace nb200
