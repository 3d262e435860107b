// Device building blocks shared by the bound kernels.
//
// Canonical arithmetic (the order oracle/c/nb200_oracle.c restates, so fp64
// values can be compared bit-for-bit):
//   whitening  t_i = sum_j M[i][j] * s_j      left-to-right FMA chain from 0
//   radius     r2  = sum_i t_i * t_i          left-to-right FMA chain from 0
//   MLP layer  h_n = (sum_k W[k][n] * a_k) + b_n, ReLU = max(h, 0)
//   ensemble   ((p_0 + p_1) + ...) / n_net
// Rows of points live in shared memory, one row per thread, stride d|1 so
// that 64-bit accesses of a warp are bank-conflict free.
#pragma once
#include "nb200_common.cuh"

namespace nb200 {

// Coalesced copy of rows [base, base + nrows) of a row-major [n, d] array
// into per-thread shared rows (and back).
__device__ __forceinline__ void load_rows(const double* __restrict__ src,
                                          int64_t base, int nrows, int d,
                                          int stride, double* sm) {
  const double* g = src + base * (int64_t)d;
  const int total = nrows * d;
  // (row, col) advance incrementally: no integer division per element
  const int step_r = (int)blockDim.x / d, step_c = (int)blockDim.x % d;
  int r = (int)threadIdx.x / d, c = (int)threadIdx.x % d;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    sm[r * stride + c] = g[e];
    r += step_r; c += step_c;
    if (c >= d) { c -= d; r += 1; }
  }
}
__device__ __forceinline__ void store_rows(double* __restrict__ dst,
                                           int64_t base, int nrows, int d,
                                           int stride, const double* sm) {
  double* g = dst + base * (int64_t)d;
  const int total = nrows * d;
  const int step_r = (int)blockDim.x / d, step_c = (int)blockDim.x % d;
  int r = (int)threadIdx.x / d, c = (int)threadIdx.x % d;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    g[e] = sm[r * stride + c];
    r += step_r; c += step_c;
    if (c >= d) { c -= d; r += 1; }
  }
}

// out_i = sum_j M[i][j] * s[j] for i < de, handed to f(i, value) in
// increasing i.  `lower` != 0 promises M's strict upper triangle is exactly
// zero, in which case the skipped terms are exact no-ops (fma(0, s, a) == a).
// Four rows share each s[j] load; every row keeps its own left-to-right chain.
template <typename F>
__device__ __forceinline__ void matvec_rows(const double* __restrict__ M,
                                            int de, int lower,
                                            const double* s, F&& f) {
  int i = 0;
  for (; i + 4 <= de; i += 4) {
    const double* m0 = M + (size_t)i * de;
    const double* m1 = m0 + de;
    const double* m2 = m1 + de;
    const double* m3 = m2 + de;
    const int jmax = lower ? i + 4 : de;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int j = 0; j < jmax; ++j) {
      const double sj = s[j];
      a0 = fma(__ldg(m0 + j), sj, a0);
      a1 = fma(__ldg(m1 + j), sj, a1);
      a2 = fma(__ldg(m2 + j), sj, a2);
      a3 = fma(__ldg(m3 + j), sj, a3);
    }
    f(i, a0); f(i + 1, a1); f(i + 2, a2); f(i + 3, a3);
  }
  for (; i < de; ++i) {
    const double* m0 = M + (size_t)i * de;
    const int jmax = lower ? i + 1 : de;
    double a0 = 0.0;
    for (int j = 0; j < jmax; ++j) a0 = fma(__ldg(m0 + j), s[j], a0);
    f(i, a0);
  }
}

// Squared Mahalanobis radius of the row x restricted to dims idx[0..de)
// (idx == nullptr: identity).  s is a scratch row (>= de); t (optional, may
// not alias s) receives the whitened coordinates.
__device__ __forceinline__ double whiten_r2(const double* x,
                                            const int32_t* __restrict__ idx,
                                            int de,
                                            const double* __restrict__ c,
                                            const double* __restrict__ Binv,
                                            int lower, double* s, double* t) {
  if (idx) {
    for (int j = 0; j < de; ++j) s[j] = x[__ldg(idx + j)] - __ldg(c + j);
  } else {
    for (int j = 0; j < de; ++j) s[j] = x[j] - __ldg(c + j);
  }
  double r2 = 0.0;
  if (t) {
    matvec_rows(Binv, de, lower, s, [&](int i, double v) {
      t[i] = v;
      r2 = fma(v, v, r2);
    });
  } else {
    matvec_rows(Binv, de, lower, s,
                [&](int, double v) { r2 = fma(v, v, r2); });
  }
  return r2;
}

// UnitCube.contains on selected dims (nautilus/bounds/basic.py:67).
__device__ __forceinline__ bool cube_ok(const double* x,
                                        const int32_t* __restrict__ idx,
                                        int n) {
  bool ok = true;
  if (idx) {
    for (int j = 0; j < n; ++j) {
      const double v = x[__ldg(idx + j)];
      ok = ok && (v >= 0.0) && (v < 1.0);
    }
  } else {
    for (int j = 0; j < n; ++j) {
      const double v = x[j];
      ok = ok && (v >= 0.0) && (v < 1.0);
    }
  }
  return ok;
}

// UnitCubeEllipsoidMixture.contains (nautilus/bounds/basic.py:610-617).
__device__ __forceinline__ bool mix_contains(const Rec& rec,
                                             const double* __restrict__ data,
                                             int k, const double* x,
                                             double* s) {
  const int32_t* m = rec.mix(k);
  const int de = m[0], nc = m[1];
  const int32_t* idx = rec.r + m[2];
  bool in = true;
  if (nc > 0) in = cube_ok(x, idx + de, nc);
  if (de > 0) {
    const double r2 = whiten_r2(x, nc > 0 ? idx : nullptr, de, data + m[3],
                                data + m[5], m[6], s, nullptr);
    in = in && (r2 < 1.0);
  }
  return in;
}

// sum_k contains_k (nautilus/bounds/union.py:316-317).
__device__ __forceinline__ int union_count(const Rec& rec,
                                           const double* __restrict__ data,
                                           const double* x, double* s) {
  int cnt = 0;
  const int K = rec.K();
  for (int k = 0; k < K; ++k) cnt += mix_contains(rec, data, k, x, s) ? 1 : 0;
  return cnt;
}

// ---- online log-sum-exp triple -------------------------------------------
struct Lse {
  double m, s1, s2;  // max, sum exp(l-m), sum exp(2(l-m))
  __device__ __forceinline__ void init() {
    m = -INFINITY; s1 = 0.0; s2 = 0.0;
  }
  __device__ __forceinline__ void add(double l) {
    if (!(l > -INFINITY)) return;  // -inf contributes 0; NaN is skipped
    if (l > m) {
      const double e = exp(m - l);  // exp(-inf) = 0 on first element
      s1 = fma(s1, e, 1.0);
      s2 = fma(s2, e * e, 1.0);
      m = l;
    } else {
      const double e = exp(l - m);
      s1 += e;
      s2 = fma(e, e, s2);
    }
  }
  __device__ __forceinline__ void merge(const Lse& o) {
    if (!(o.m > -INFINITY)) return;
    if (!(m > -INFINITY)) { *this = o; return; }
    if (o.m > m) {
      const double e = exp(m - o.m);
      s1 = fma(s1, e, o.s1);
      s2 = fma(s2, e * e, o.s2);
      m = o.m;
    } else {
      const double e = exp(o.m - m);
      s1 = fma(o.s1, e, s1);
      s2 = fma(o.s2, e * e, s2);
    }
  }
};


// ---- built-in likelihoods; x may live in any memory space --
__device__ __forceinline__ double loglike_eval(int like_id,
                                               const double* __restrict__ p,
                                               const double* x, int d) {
  double out;
  if (like_id == NB200_LIKE_GAUSSIAN) {
    const double* mu = p + 2;
    double s2 = 0.0;
    // loads are issued eight at a time (x may be a global row), the sum keeps
    // its left-to-right FMA order
    int j = 0;
    for (; j + 8 <= d; j += 8) {
      double v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = x[j + q];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] -= __ldg(mu + j + q);
#pragma unroll
      for (int q = 0; q < 8; ++q) s2 = fma(v[q], v[q], s2);
    }
    for (; j < d; ++j) {
      const double v = x[j] - __ldg(mu + j);
      s2 = fma(v, v, s2);
    }
    out = -0.5 * __ldg(p) * s2 + __ldg(p + 1);
  } else if (like_id == NB200_LIKE_ROSENBROCK) {
    const double lo = __ldg(p), w = __ldg(p + 1);
    double acc = 0.0;
    double cur = lo + w * x[0];
    for (int j = 0; j + 1 < d; ++j) {
      const double nxt = lo + w * x[j + 1];
      const double a = nxt - cur * cur;
      const double b = 1.0 - cur;
      acc += 100.0 * a * a + b * b;
      cur = nxt;
    }
    out = -acc;
  } else if (like_id == NB200_LIKE_MIXTURE) {
    const int M = (int)__ldg(p);
    const double is2 = __ldg(p + 1), norm = __ldg(p + 2);
    const double* mu = p + 3;
    Lse acc;
    acc.init();
    for (int mth = 0; mth < M; ++mth) {
      double s2 = 0.0;
      for (int j = 0; j < d; ++j) {
        const double v = x[j] - __ldg(mu + mth * d + j);
        s2 = fma(v, v, s2);
      }
      acc.add(-0.5 * is2 * s2);
    }
    out = acc.m + log(acc.s1) - log((double)M) + norm;
  } else {  // NB200_LIKE_EQUICORR
    const double a = __ldg(p), b = __ldg(p + 1), norm = __ldg(p + 2);
    const double* mu = p + 3;
    double s1 = 0.0, s2 = 0.0;
    for (int j = 0; j < d; ++j) {
      const double v = x[j] - __ldg(mu + j);
      s1 += v;
      s2 = fma(v, v, s2);
    }
    out = -0.5 * (a * s2 - b * s1 * s1) + norm;
  }
  return out;
}

// ---- shell sums: per-block partials -----------------------------------------
constexpr int STAT_THREADS = 256;
constexpr int STAT_MAX_BLOCKS = 1184;  // 148 SMs x 8 resident CTAs

struct StatPartial {
  double m, s1, s2, pad;
  long long cnt[NB200_N_CNT];
};

// Deterministic block reduction (warp shuffles, then warp leaders in order).
template <int THREADS>
__device__ __forceinline__ void stat_block_reduce(Lse& acc, long long* cnt,
                                                  StatPartial* dst) {
  __shared__ double sm_m[THREADS / 32], sm_s1[THREADS / 32],
      sm_s2[THREADS / 32];
  __shared__ long long sm_c[THREADS / 32][NB200_N_CNT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Lse other;
    other.m = __shfl_down_sync(0xffffffffu, acc.m, o);
    other.s1 = __shfl_down_sync(0xffffffffu, acc.s1, o);
    other.s2 = __shfl_down_sync(0xffffffffu, acc.s2, o);
    acc.merge(other);
#pragma unroll
    for (int q = 0; q < NB200_N_CNT; ++q)
      cnt[q] += __shfl_down_sync(0xffffffffu, cnt[q], o);
  }
  if (lane == 0) {
    sm_m[warp] = acc.m; sm_s1[warp] = acc.s1; sm_s2[warp] = acc.s2;
#pragma unroll
    for (int q = 0; q < NB200_N_CNT; ++q) sm_c[warp][q] = cnt[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Lse tot;
    tot.m = sm_m[0]; tot.s1 = sm_s1[0]; tot.s2 = sm_s2[0];
    long long c[NB200_N_CNT];
#pragma unroll
    for (int q = 0; q < NB200_N_CNT; ++q) c[q] = sm_c[0][q];
    for (int w = 1; w < THREADS / 32; ++w) {
      Lse o;
      o.m = sm_m[w]; o.s1 = sm_s1[w]; o.s2 = sm_s2[w];
      tot.merge(o);
#pragma unroll
      for (int q = 0; q < NB200_N_CNT; ++q) c[q] += sm_c[w][q];
    }
    dst->m = tot.m; dst->s1 = tot.s1; dst->s2 = tot.s2; dst->pad = 0.0;
#pragma unroll
    for (int q = 0; q < NB200_N_CNT; ++q) dst->cnt[q] = c[q];
  }
}

}  // namespace nb200
