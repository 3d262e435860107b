// Two-component Gaussian mixture by EM, all restarts in ONE launch
// ("GMM-2 split (EM, n_init=10)").
//
// The reference splits a bound with scikit-learn's
// GaussianMixture(n_components=2, n_init=10) on the whitened points
// (nautilus/bounds/union.py:185-187): ten restarts of a strictly sequential
// EM iteration on a few thousand points in a few tens of dimensions --
// ~1 MFLOP per step, latency, not throughput.  The host restatement
// (nautilus_b200/bounds/_construct.py: two_gaussians) is the algorithm; the
// batched tensor form of it (two_gaussians_batched) costs ~45 library
// launches per EM iteration.  Here one thread-block CLUSTER runs one restart
// start to finish:
//
//   * the points are split over the CTAs of the cluster and stay in shared
//     memory (or L2 when they do not fit) for the whole fit;
//   * M step: every CTA accumulates the weighted moments of BOTH components
//     over its points -- S~ = sum_p r_pk x~_p x~_p^T with x~ = (x, 1), i.e.
//     S2, S1 and S0 in one packed triangle -- and the cluster all-reduces
//     them through distributed shared memory in a fixed order, so every CTA
//     holds bit-identical totals and takes identical decisions (no
//     broadcast); double-buffered: ONE cluster barrier per EM iteration;
//   * every CTA factors both covariances (Cholesky, one warp each) and
//     inverts the factors (W = L^-1), redundantly -- cheaper than a second
//     exchange;
//   * E step: one thread per (point, component): z = W x - W mu as a
//     triangular matrix-vector product out of shared memory, log joint
//     density, responsibilities of the next iteration; the log densities of
//     the last accepted iteration stay in global memory.
//
// Stopping / abandoning rules are those of two_gaussians: a restart stops
// when the mean log likelihood moves by less than tol (its score is the value
// BEFORE the last step), when a component starves (fewer than d + 1 points:
// keeps what it had) or when a covariance is not positive definite
// (discarded).  fp64 throughout.
#include <cooperative_groups.h>
#include <math_constants.h>

#include "nb200_common.cuh"

namespace cg = cooperative_groups;

namespace nb200 {

constexpr int GM_THREADS = 512;
constexpr int GM_WARPS = GM_THREADS / 32;
constexpr int GM_CLUSTER = 8;
constexpr int GM_D_MAX = 64;

struct GmArgs {
  const double* x;        // [n, d] row-major
  const uint8_t* label;   // [R, n] initial hard assignment (0 / 1)
  int n, d, max_iter;
  int ld;                 // points per CTA
  int x_resident;         // points copied to shared memory
  double tol, reg;
  double* log_p;          // [R, 2, n]
  double* score;          // [R]
  int32_t* iters;         // [R]
};

// doubles of shared memory (0: the shape does not fit at all)
__host__ __device__ inline size_t gm_smem_doubles(int ld, int d, int resident) {
  const int D = d + 1, T = D * (D + 1) / 2, NS = 2 * T + 2;
  return (size_t)(resident ? ld * (size_t)(d | 1) : 0) + 2 * (size_t)ld +
         3 * (size_t)NS + 4 * (size_t)d * d + 4 * (size_t)d + 64;
}

__global__ void __launch_bounds__(GM_THREADS, 1)
k_gmm2_em(const GmArgs A) {
  extern __shared__ __align__(16) double sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int restart = blockIdx.x / C;
  const int n = A.n, d = A.d, D = d + 1;
  const int T = D * (D + 1) / 2;          // packed triangle of x~ x~^T
  const int NS = 2 * T + 2;               // both components + ll (+ pad)
  const int LL = 2 * T;
  const int ld = A.ld, S = d | 1;
  const int p0 = rank * ld;
  const int np = max(0, min(ld, n - p0));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  double* xs = sm;                                   // ld x S (if resident)
  double* resp = xs + (A.x_resident ? (size_t)ld * S : 0);   // [2][ld]
  double* stat = resp + 2 * (size_t)ld;              // [2 buffers][NS]
  double* tot = stat + 2 * (size_t)NS;               // [NS]
  double* Lm = tot + NS;                             // [2][d*d]
  double* W = Lm + 2 * (size_t)d * d;                // [2][d*d]
  double* mean = W + 2 * (size_t)d * d;              // [2][d]
  double* bvec = mean + 2 * d;                       // [2][d]: W mu
  double* misc = bvec + 2 * d;                       // [64]
  // misc: 0,1 logdet; 2,3 log(nk / n); 4 fail flag; 8.. reduction scratch

  const double* xb;
  int xst;
  if (A.x_resident) {
    for (int e = tid; e < np * d; e += GM_THREADS) {
      const int p = e / d, j = e - p * d;
      xs[p * S + j] = A.x[(size_t)(p0 + p) * d + j];
    }
    xb = xs; xst = S;
  } else {
    xb = A.x + (size_t)p0 * d; xst = d;
  }
  for (int p = tid; p < ld; p += GM_THREADS) {
    const double one = p < np ? (double)A.label[(size_t)restart * n + p0 + p]
                              : 0.0;
    resp[p] = p < np ? 1.0 - one : 0.0;
    resp[ld + p] = one;
  }
  if (tid == 0) misc[5] = 0.0;             // ll of the last E step (local)
  __syncthreads();

  int it = 0, valid = 0;
  double ll_old = -CUDART_INF;
  const double log_2pi_term = 0.5 * d * 1.8378770664093453;

  while (true) {
    const int buf = it & 1;
    double* my = stat + (size_t)buf * NS;
    // ---- M step, local part: packed moments of both components -------------
    for (int item = tid; item < T; item += GM_THREADS) {
      // item -> (i, j), i >= j, row-major packed lower triangle
      int i = (int)((sqrt(8.0 * item + 1.0) - 1.0) * 0.5);
      while ((i + 1) * (i + 2) / 2 <= item) ++i;
      while (i * (i + 1) / 2 > item) --i;
      const int j = item - i * (i + 1) / 2;
      const bool io = i == d, jo = j == d;
      double a0 = 0.0, a1 = 0.0;
      for (int p = 0; p < np; ++p) {
        const double xi = io ? 1.0 : xb[(size_t)p * xst + i];
        const double xj = jo ? 1.0 : xb[(size_t)p * xst + j];
        const double pr = xi * xj;
        a0 = fma(resp[p], pr, a0);
        a1 = fma(resp[ld + p], pr, a1);
      }
      my[item] = a0;
      my[T + item] = a1;
    }
    if (tid == 0) { my[LL] = misc[5]; my[LL + 1] = 0.0; }
    // ---- all-reduce over the cluster, fixed order --------------------------
    cluster.sync();
    for (int t = tid; t < NS; t += GM_THREADS) {
      double acc = 0.0;
      for (int r = 0; r < C; ++r)
        acc += cluster.map_shared_rank(stat, r)[(size_t)buf * NS + t];
      tot[t] = acc;
    }
    __syncthreads();
    // ---- decisions (identical in every thread of every CTA) ----------------
    if (it > 0) {
      const double ll = tot[LL] / (double)n;
      if (fabs(ll - ll_old) < A.tol) break;
      ll_old = ll;
    }
    if (it >= A.max_iter) break;
    const double s0a = tot[T - 1], s0b = tot[2 * T - 1];   // (i, j) = (d, d)
    const double nka = s0a + 1e-12, nkb = s0b + 1e-12;
    if (nka < D || nkb < D) break;                          // starved
    // ---- parameters: mean, covariance -> Lm ---------------------------------
    const int base1 = d * (d + 1) / 2;      // first item of row d: S1
    for (int e = tid; e < 2 * d; e += GM_THREADS) {
      const int k = e / d, i = e - k * d;
      mean[e] = tot[k * T + base1 + i] / (k ? nkb : nka);
    }
    if (tid < 2) {
      misc[2 + tid] = log((tid ? nkb : nka) / (double)n);
      if (tid == 0) misc[4] = 0.0;
    }
    __syncthreads();
    for (int e = tid; e < 2 * base1; e += GM_THREADS) {
      const int k = e / base1, item = e - k * base1;
      int i = (int)((sqrt(8.0 * item + 1.0) - 1.0) * 0.5);
      while ((i + 1) * (i + 2) / 2 <= item) ++i;
      while (i * (i + 1) / 2 > item) --i;
      const int j = item - i * (i + 1) / 2;
      const double* tk = tot + k * T;
      const double nk = k ? nkb : nka, s0 = k ? s0b : s0a;
      const double mi = mean[k * d + i], mj = mean[k * d + j];
      // sum r (x - m)(x - m)^T = S2 - m S1^T - S1 m^T + S0 m m^T
      double c = tk[item] - mi * tk[base1 + j] - tk[base1 + i] * mj +
                 s0 * mi * mj;
      c = c / nk + (i == j ? A.reg : 0.0);
      Lm[(size_t)k * d * d + i * d + j] = c;
    }
    __syncthreads();
    // ---- Cholesky (left-looking, one warp per component) and W = L^-1 ------
    if (warp < 2) {
      double* L = Lm + (size_t)warp * d * d;
      double* Wk = W + (size_t)warp * d * d;
      double logdet = 0.0;
      bool ok = true;
      for (int j = 0; j < d; ++j) {
        // column j: rows i >= j
        for (int i = j + lane; i < d; i += 32) {
          double v = L[i * d + j];
          for (int c = 0; c < j; ++c) v = fma(-L[i * d + c], L[j * d + c], v);
          L[i * d + j] = v;          // unscaled; scaled below
        }
        __syncwarp();
        const double diag = L[j * d + j];
        if (!(diag > 0.0) || !isfinite(diag)) { ok = false; break; }
        const double ljj = sqrt(diag);
        logdet += log(ljj);
        for (int i = j + lane; i < d; i += 32)
          L[i * d + j] = i == j ? ljj : L[i * d + j] / ljj;
        __syncwarp();
      }
      if (!ok) {
        if (lane == 0) misc[4] = 1.0;
      } else {
        if (lane == 0) misc[warp] = logdet;
        // W = L^-1, row by row; lane owns column c
        for (int i = 0; i < d; ++i) {
          const double inv = 1.0 / L[i * d + i];
          for (int c = lane; c <= i; c += 32) {
            double acc = 0.0;
            for (int m = c; m < i; ++m)
              acc = fma(L[i * d + m], Wk[m * d + c], acc);
            Wk[i * d + c] = c == i ? inv : -acc * inv;
          }
          __syncwarp();
        }
        // b = W mu
        for (int i = lane; i < d; i += 32) {
          double acc = 0.0;
          for (int j = 0; j <= i; ++j)
            acc = fma(Wk[i * d + j], mean[warp * d + j], acc);
          bvec[warp * d + i] = acc;
        }
      }
    }
    __syncthreads();
    if (misc[4] != 0.0) { valid = 0; break; }       // not positive definite
    // ---- E step: thread per (point, component) ------------------------------
    double ll_part = 0.0;
    const int items = ((2 * ld + 31) / 32) * 32;
    for (int e = tid; e < items; e += GM_THREADS) {
      const int p = e >> 1, k = e & 1;
      double lp = -CUDART_INF;
      if (p < np) {
        const double* Wk = W + (size_t)k * d * d;
        const double* xr = xb + (size_t)p * xst;
        const double* bk = bvec + k * d;
        double maha = 0.0;
        for (int i = 0; i < d; ++i) {
          double z = -bk[i];
          for (int j = 0; j <= i; ++j) z = fma(Wk[i * d + j], xr[j], z);
          maha = fma(z, z, maha);
        }
        lp = -0.5 * maha - misc[k] - log_2pi_term + misc[2 + k];
      }
      const double other = __shfl_xor_sync(0xffffffffu, lp, 1);
      if (p < np) {
        const double m = fmax(lp, other);
        const double norm = m + log(exp(lp - m) + exp(other - m));
        resp[k * ld + p] = exp(lp - norm);
        A.log_p[((size_t)restart * 2 + k) * n + p0 + p] = lp;
        if (k == 0) ll_part += norm;
      }
    }
    ll_part = warp_sum(ll_part);
    if (lane == 0) misc[8 + warp] = ll_part;
    __syncthreads();
    if (tid == 0) {
      double acc = 0.0;
      for (int w = 0; w < GM_WARPS; ++w) acc += misc[8 + w];
      misc[5] = acc;
    }
    __syncthreads();
    valid = 1;
    ++it;
  }
  // peers may still be reading this CTA's partial sums
  cluster.sync();
  if (rank == 0 && tid == 0) {
    A.score[restart] = valid ? ll_old : -CUDART_INF;
    A.iters[restart] = it;
  }
}

static int gm_plan(int64_t n, int d, int* ld, int* resident, size_t* smem) {
  if (d < 1 || d > GM_D_MAX || n < 2 || n >= (1ll << 30)) return 0;
  *ld = (int)((n + GM_CLUSTER - 1) / GM_CLUSTER);
  for (int res = 1; res >= 0; --res) {
    const size_t bytes = 8 * gm_smem_doubles(*ld, d, res);
    if (bytes <= 200 * 1024) {
      *resident = res; *smem = bytes;
      return 1;
    }
  }
  return 0;
}

}  // namespace nb200

using namespace nb200;

extern "C" {

int nb200_gmm2_applicable(int64_t n, int d) {
  int ld, res; size_t smem;
  return gm_plan(n, d, &ld, &res, &smem);
}

int nb200_gmm2_em(const double* x_d, int64_t n, int d,
                  const uint8_t* label_d, int n_init, int max_iter, double tol,
                  double reg, double* log_p_d, double* score_d,
                  int32_t* iters_d, void* stream) {
  int ld = 0, res = 0; size_t smem = 0;
  NB_CHECK(gm_plan(n, d, &ld, &res, &smem),
           "gmm2_em: shape outside the kernel's envelope (d <= 64, points "
           "per CTA must fit shared memory)");
  NB_CHECK(n_init >= 1 && n_init <= 4096 && max_iter >= 1 && tol > 0.0 &&
           reg >= 0.0, "gmm2_em: bad iteration controls");
  GmArgs A;
  A.x = x_d; A.label = label_d; A.n = (int)n; A.d = d; A.max_iter = max_iter;
  A.ld = ld; A.x_resident = res; A.tol = tol; A.reg = reg;
  A.log_p = log_p_d; A.score = score_d; A.iters = iters_d;
  NB_CUDA(cudaFuncSetAttribute(k_gmm2_em,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(GM_CLUSTER * n_init, 1, 1);
  cfg.blockDim = dim3(GM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = GM_CLUSTER; attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  ProfScope prof(ST_FIT, (cudaStream_t)stream);
  NB_CUDA(cudaLaunchKernelEx(&cfg, k_gmm2_em, A));
  NB_LAUNCH_OK();
  return 0;
}

}  // extern "C"
