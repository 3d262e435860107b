// Bound construction on the device.
//
// nb200_mvee_weights: Khachiyan's first-order algorithm for the
// minimum-volume enclosing ellipsoid in the lifted space (Todd & Yildirim
// 2007), the job of minimum_volume_enclosing_ellipsoid in the reference
// (nautilus/bounds/basic.py:175-241, called ~130 times per config-2 run from
// Ellipsoid.compute :266-316 and UnitCubeEllipsoidMixture.compute :471-563).
// Same iteration as the host restatement (nautilus_b200/bounds/_construct.py:
// _khachiyan): Sherman-Morrison rank-one update of the inverse moment matrix,
// O(N d) update of the Mahalanobis distances, a from-scratch refresh now and
// then (MV_REFRESH).
//
// The loop is a few thousand strictly sequential steps of a few hundred
// kFLOP: latency, not throughput.  ONE persistent CTA of 1024 threads runs it
// start to finish: the (d+1) x (d+1) inverse lives in shared memory, the
// distances g and the weights u in L2-resident global memory, the points are
// read coordinate-major (qT[k][i]) so that every pass over the points is
// coalesced.  fp64 throughout.
#include <cooperative_groups.h>

#include "nb200_common.cuh"

namespace cg = cooperative_groups;

namespace nb200 {

constexpr int MV_THREADS = 1024;
constexpr int MV_WARPS = MV_THREADS / 32;
// rank-one updates between two from-scratch refreshes of the inverse (the
// host restatement refreshes every 64; the drift of 512 Sherman-Morrison
// updates on whitened points is ~1e-12, far below the 1e-3 stopping rule)
constexpr int MV_REFRESH = 512;

struct MvShared {
  double red_v[MV_WARPS];
  int red_i[MV_WARPS];
  double gmax;
  int jmax;
  int stop;
};

// (value, index) max with ties to the smaller index (numpy.argmax)
__device__ __forceinline__ void arg_better(double& v, int& i, double ov,
                                           int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

__device__ __forceinline__ void block_argmax(double v, int i, MvShared* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_down_sync(0xffffffffu, v, o);
    const int oi = __shfl_down_sync(0xffffffffu, i, o);
    arg_better(v, i, ov, oi);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh->red_v[warp] = v; sh->red_i[warp] = i; }
  __syncthreads();
  if (warp == 0) {
    v = sh->red_v[lane];
    i = sh->red_i[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, v, o);
      const int oi = __shfl_down_sync(0xffffffffu, i, o);
      arg_better(v, i, ov, oi);
    }
    if (lane == 0) { sh->gmax = v; sh->jmax = i; }
  }
  __syncthreads();
}

// lifted coordinate k of point i (k == d: the constant 1)
__device__ __forceinline__ double lifted(const double* __restrict__ qT,
                                         int64_t n, int d, int k, int64_t i) {
  return k < d ? qT[(int64_t)k * n + i] : 1.0;
}

__global__ void __launch_bounds__(MV_THREADS, 1)
k_mvee(const double* __restrict__ qT, int64_t n, int d, int max_updates,
       double tol, double* __restrict__ u, double* __restrict__ g,
       int32_t* __restrict__ iters_out) {
  extern __shared__ __align__(16) double smv[];
  __shared__ MvShared sh;
  const int D = d + 1;
  const int Dp = D | 1;                 // odd row stride: conflict-free columns
  double* Vinv = smv;                   // [D][Dp]
  double* w = Vinv + (size_t)D * Dp;    // [D]
  double* qj = w + D;                   // [D]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int64_t i = tid; i < n; i += MV_THREADS) u[i] = 1.0 / (double)n;
  __syncthreads();

  // V = sum_i u_i q_i q_i^T (one warp per entry of the upper triangle, lanes
  // over the points), inverse by in-place Gauss-Jordan (V is SPD), then
  // g_i = q_i^T V^-1 q_i; returns this thread's running (max g, argmax)
  auto refresh = [&](double& best_v, int& best_i) {
    for (int e = warp; e < D * D; e += MV_WARPS) {
      const int a = e / D, b = e - a * D;
      if (b < a) continue;
      double acc = 0.0;
      for (int64_t i = lane; i < n; i += 32)
        acc = fma(u[i] * lifted(qT, n, d, a, i), lifted(qT, n, d, b, i), acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_down_sync(0xffffffffu, acc, o);
      if (lane == 0) { Vinv[a * Dp + b] = acc; Vinv[b * Dp + a] = acc; }
    }
    __syncthreads();
    for (int k = 0; k < D; ++k) {
      const double pivot = 1.0 / Vinv[k * Dp + k];
      __syncthreads();
      for (int b = tid; b < D; b += MV_THREADS)
        if (b != k) Vinv[k * Dp + b] *= pivot;
      __syncthreads();
      for (int e = tid; e < D * D; e += MV_THREADS) {
        const int a = e / D, b = e - a * D;
        if (a != k && b != k)
          Vinv[a * Dp + b] =
              fma(-Vinv[a * Dp + k], Vinv[k * Dp + b], Vinv[a * Dp + b]);
      }
      __syncthreads();
      for (int a = tid; a < D; a += MV_THREADS)
        Vinv[a * Dp + k] = a == k ? pivot : -Vinv[a * Dp + k] * pivot;
      __syncthreads();
    }
    // g_i = sum_a q_a V^-1[a][a] q_a + 2 sum_{b<a} q_a V^-1[a][b] q_b, four
    // independent chains per row so that the (L1-resident) loads of q do not
    // serialise on one accumulator
    best_v = -1.0; best_i = 0x7fffffff;
    for (int64_t i = tid; i < n; i += MV_THREADS) {
      double gi = 0.0;
      for (int a = 0; a < D; ++a) {
        const double* row = Vinv + a * Dp;
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        int b = 0;
        for (; b + 4 <= a; b += 4) {
          t0 = fma(row[b], lifted(qT, n, d, b, i), t0);
          t1 = fma(row[b + 1], lifted(qT, n, d, b + 1, i), t1);
          t2 = fma(row[b + 2], lifted(qT, n, d, b + 2, i), t2);
          t3 = fma(row[b + 3], lifted(qT, n, d, b + 3, i), t3);
        }
        for (; b < a; ++b) t0 = fma(row[b], lifted(qT, n, d, b, i), t0);
        const double qa = lifted(qT, n, d, a, i);
        const double off = (t0 + t1) + (t2 + t3);
        gi = fma(qa, fma(row[a], qa, 2.0 * off), gi);
      }
      g[i] = gi;
      arg_better(best_v, best_i, gi, (int)i);
    }
  };

  double best_v; int best_i;
  refresh(best_v, best_i);
  int it = 0;
  for (; it < max_updates; ++it) {
    block_argmax(best_v, best_i, &sh);
    const double gmax = sh.gmax;
    const int j = sh.jmax;
    if (gmax <= (double)D * (1.0 + tol)) break;
    const double step = (gmax - D) / ((double)D * (gmax - 1.0));
    const double beta = step / (1.0 - step);
    const double one_m = 1.0 - step;
    if (it % MV_REFRESH == MV_REFRESH - 1) {
      // refresh the inverse from scratch now and then (rounding drift)
      for (int64_t i = tid; i < n; i += MV_THREADS)
        u[i] = u[i] * one_m + (i == j ? step : 0.0);
      __syncthreads();
      refresh(best_v, best_i);
      continue;
    }
    if (tid < D) qj[tid] = lifted(qT, n, d, tid, j);
    __syncthreads();
    if (tid < D) {
      double acc = 0.0;
      for (int b = 0; b < D; ++b) acc = fma(Vinv[tid * Dp + b], qj[b], acc);
      w[tid] = acc;
    }
    __syncthreads();
    const double coef = beta / (1.0 + beta * gmax);
    const double inv1 = 1.0 / one_m;
    // inverse and distances after V <- (1 - step) V + step q_j q_j^T
    for (int e = tid; e < D * D; e += MV_THREADS) {
      const int a = e / D, b = e - a * D;
      Vinv[a * Dp + b] = (Vinv[a * Dp + b] - coef * w[a] * w[b]) * inv1;
    }
    best_v = -1.0; best_i = 0x7fffffff;
    for (int64_t i = tid; i < n; i += MV_THREADS) {
      // two chains: the coalesced loads of q stay in flight
      double dot = w[d], dot1 = 0.0;
      int k = 0;
#pragma unroll 4
      for (; k + 2 <= d; k += 2) {
        dot = fma(qT[(int64_t)k * n + i], w[k], dot);
        dot1 = fma(qT[(int64_t)(k + 1) * n + i], w[k + 1], dot1);
      }
      if (k < d) dot = fma(qT[(int64_t)k * n + i], w[k], dot);
      dot += dot1;
      const double gi = (g[i] - coef * dot * dot) * inv1;
      g[i] = gi;
      u[i] = u[i] * one_m + (i == j ? step : 0.0);
      arg_better(best_v, best_i, gi, (int)i);
    }
    __syncthreads();
  }
  if (tid == 0 && iters_out) *iters_out = it;
}


// ---------------------------------------------------------------------------
// Cluster variant: the points do not fit one CTA's shared memory (2 000 x 30
// doubles = 480 KB) and re-reading them from L2 on every rank-one update is
// what bounds k_mvee (one SM pulls ~64 B/clk out of L2).  Here a thread-block
// cluster of C CTAs keeps the points RESIDENT: CTA r owns the contiguous
// slice [r n/C, (r+1) n/C) -- coordinates, distances g and weights u -- in its
// own shared memory; every CTA carries its own copy of the inverse moment
// matrix and applies the same update to it (bit-identical: the same
// arithmetic on the same operands).  Per update the CTAs exchange, through
// distributed shared memory, one candidate each -- (max g, argmax) AND that
// point's lifted coordinates -- so one cluster barrier per update decides the
// winner and delivers its coordinates.
// ---------------------------------------------------------------------------
constexpr int MVC_THREADS = 512;
constexpr int MVC_WARPS = MVC_THREADS / 32;
constexpr int MVC_MAX_CLUSTER = 16;   // 16 needs the non-portable opt-in

struct MvcShared {
  double red_v[MVC_WARPS];
  int red_i[MVC_WARPS];
  // written by the peers; double-buffered by the parity of the update: with
  // ONE cluster barrier per update a fast CTA writes the candidates of
  // update t + 1 while a slow one still reads those of update t
  double cand_v[2][MVC_MAX_CLUSTER];
  int cand_i[2][MVC_MAX_CLUSTER];
  int my_i;                         // this CTA's candidate (global index)
};

__global__ void __launch_bounds__(MVC_THREADS, 1)
k_mvee_cluster(const double* __restrict__ qT, int n, int d, int n_loc_max,
               int max_updates, double tol, double* __restrict__ u_out,
               int32_t* __restrict__ iters_out) {
  extern __shared__ __align__(16) double smc[];
  __shared__ MvcShared sh;
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int D = d + 1, Dp = D | 1;
  const int lo = (int)((long long)n * rank / C);
  const int hi = (int)((long long)n * (rank + 1) / C);
  const int nl = hi - lo;                 // points this CTA owns
  const int ld = n_loc_max;               // row stride of the local slices
  double* Vinv = smc;                     // [D][Dp]
  double* Vpart = Vinv + (size_t)D * Dp;  // [D][Dp]  this CTA's partial V
  double* w = Vpart + (size_t)D * Dp;     // [D]
  double* cq = w + D;                     // [2][MAX_CLUSTER][D] the candidates'
                                          // lifted coordinates (by the peers)
  double* g = cq + (size_t)2 * MVC_MAX_CLUSTER * D;   // [ld]
  double* u = g + ld;                     // [ld]
  double* q = u + ld;                     // [d][ld] coordinate-major slice
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int e = tid; e < d * nl; e += MVC_THREADS) {
    const int k = e / nl, i = e - k * nl;
    q[k * ld + i] = qT[(int64_t)k * n + lo + i];
  }
  for (int i = tid; i < nl; i += MVC_THREADS) u[i] = 1.0 / (double)n;
  __syncthreads();
  auto lift = [&](int k, int i) { return k < d ? q[k * ld + i] : 1.0; };

  // (max, argmax) of this thread's candidates over the whole cluster;
  // ties go to the smaller global index (numpy.argmax)
  auto cluster_argmax = [&](double v, int i, int par, double& gmax, int& jmax,
                            int& rwin) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, v, o);
      const int oi = __shfl_down_sync(0xffffffffu, i, o);
      arg_better(v, i, ov, oi);
    }
    if (lane == 0) { sh.red_v[warp] = v; sh.red_i[warp] = i; }
    __syncthreads();
    if (warp == 0) {
      v = lane < MVC_WARPS ? sh.red_v[lane] : -1.0;
      i = lane < MVC_WARPS ? sh.red_i[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, o);
        const int oi = __shfl_down_sync(0xffffffffu, i, o);
        arg_better(v, i, ov, oi);
      }
      // lanes 0..C-1 deliver this CTA's candidate to every CTA
      v = __shfl_sync(0xffffffffu, v, 0);
      i = __shfl_sync(0xffffffffu, i, 0);
      if (lane < C) {
        MvcShared* peer = cluster.map_shared_rank(&sh, lane);
        peer->cand_v[par][rank] = v;
        peer->cand_i[par][rank] = i;
      }
      if (lane == 0) sh.my_i = i;
    }
    __syncthreads();
    {
      // ... and the candidate's lifted coordinates, to every CTA
      const int bi = sh.my_i;
      if (tid < D && bi >= lo && bi < hi) {
        const double c = lift(tid, bi - lo);
        for (int r = 0; r < C; ++r)
          cluster.map_shared_rank(cq, r)[(par * MVC_MAX_CLUSTER + rank) * D +
                                         tid] = c;
      }
    }
    cluster.sync();
    gmax = sh.cand_v[par][0]; jmax = sh.cand_i[par][0]; rwin = 0;
    for (int r = 1; r < C; ++r) {
      const int before = jmax;
      arg_better(gmax, jmax, sh.cand_v[par][r], sh.cand_i[par][r]);
      if (jmax != before) rwin = r;
    }
  };

  // V = sum_i u_i q_i q_i^T: partial sums per CTA (one warp per entry of the
  // upper triangle, lanes over the local points), summed over the cluster in
  // rank order, inverted by every CTA; then g_i = q_i^T V^-1 q_i
  auto refresh = [&](double& best_v, int& best_i) {
    for (int e = warp; e < D * D; e += MVC_WARPS) {
      const int a = e / D, b = e - a * D;
      if (b < a) continue;
      double acc = 0.0;
      for (int i = lane; i < nl; i += 32)
        acc = fma(u[i] * lift(a, i), lift(b, i), acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_down_sync(0xffffffffu, acc, o);
      if (lane == 0) { Vpart[a * Dp + b] = acc; Vpart[b * Dp + a] = acc; }
    }
    cluster.sync();
    for (int e = tid; e < D * D; e += MVC_THREADS) {
      const int a = e / D, b = e - a * D;
      double acc = 0.0;
      for (int r = 0; r < C; ++r)
        acc += cluster.map_shared_rank(Vpart, r)[a * Dp + b];
      Vinv[a * Dp + b] = acc;
    }
    cluster.sync();          // peers are done reading this CTA's partial
    for (int k = 0; k < D; ++k) {
      const double pivot = 1.0 / Vinv[k * Dp + k];
      __syncthreads();
      for (int b = tid; b < D; b += MVC_THREADS)
        if (b != k) Vinv[k * Dp + b] *= pivot;
      __syncthreads();
      for (int e = tid; e < D * D; e += MVC_THREADS) {
        const int a = e / D, b = e - a * D;
        if (a != k && b != k)
          Vinv[a * Dp + b] =
              fma(-Vinv[a * Dp + k], Vinv[k * Dp + b], Vinv[a * Dp + b]);
      }
      __syncthreads();
      for (int a = tid; a < D; a += MVC_THREADS)
        Vinv[a * Dp + k] = a == k ? pivot : -Vinv[a * Dp + k] * pivot;
      __syncthreads();
    }
    best_v = -1.0; best_i = 0x7fffffff;
    for (int i = tid; i < nl; i += MVC_THREADS) {
      double gi = 0.0;
      for (int a = 0; a < D; ++a) {
        const double* row = Vinv + a * Dp;
        double t0 = 0.0, t1 = 0.0;
        int b = 0;
        for (; b + 2 <= a; b += 2) {
          t0 = fma(row[b], lift(b, i), t0);
          t1 = fma(row[b + 1], lift(b + 1, i), t1);
        }
        if (b < a) t0 = fma(row[b], lift(b, i), t0);
        const double qa = lift(a, i);
        gi = fma(qa, fma(row[a], qa, 2.0 * (t0 + t1)), gi);
      }
      g[i] = gi;
      arg_better(best_v, best_i, gi, lo + i);
    }
  };

  double best_v; int best_i;
  refresh(best_v, best_i);
  int it = 0;
  for (; it < max_updates; ++it) {
    double gmax; int j, rwin;
    cluster_argmax(best_v, best_i, it & 1, gmax, j, rwin);
    if (gmax <= (double)D * (1.0 + tol)) break;     // same decision everywhere
    const double step = (gmax - D) / ((double)D * (gmax - 1.0));
    const double beta = step / (1.0 - step);
    const double one_m = 1.0 - step;
    if (it % MV_REFRESH == MV_REFRESH - 1) {
      for (int i = tid; i < nl; i += MVC_THREADS)
        u[i] = u[i] * one_m + (lo + i == j ? step : 0.0);
      __syncthreads();
      refresh(best_v, best_i);
      continue;
    }
    // w = V^-1 q_j: sixteen threads per row, then a shuffle tree (every CTA
    // runs the same code on the same operands: bit-identical copies)
    {
      const double* qj = cq + ((it & 1) * MVC_MAX_CLUSTER + rwin) * D;
      const int sub = tid & 15;
      for (int a0 = 0; a0 < D; a0 += MVC_THREADS / 16) {
        const int a = a0 + (tid >> 4);
        double acc = 0.0;
        if (a < D)
          for (int b = sub; b < D; b += 16)
            acc = fma(Vinv[a * Dp + b], qj[b], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (a < D && sub == 0) w[a] = acc;
      }
    }
    __syncthreads();
    const double coef = beta / (1.0 + beta * gmax);
    const double inv1 = 1.0 / one_m;
    for (int e = tid; e < D * D; e += MVC_THREADS) {
      const int a = e / D, b = e - a * D;
      Vinv[a * Dp + b] = (Vinv[a * Dp + b] - coef * w[a] * w[b]) * inv1;
    }
    best_v = -1.0; best_i = 0x7fffffff;
    for (int i = tid; i < nl; i += MVC_THREADS) {
      double dot = w[d], dot1 = 0.0;
      int k = 0;
#pragma unroll 4
      for (; k + 2 <= d; k += 2) {
        dot = fma(q[k * ld + i], w[k], dot);
        dot1 = fma(q[(k + 1) * ld + i], w[k + 1], dot1);
      }
      if (k < d) dot = fma(q[k * ld + i], w[k], dot);
      dot += dot1;
      const double gi = (g[i] - coef * dot * dot) * inv1;
      g[i] = gi;
      u[i] = u[i] * one_m + (lo + i == j ? step : 0.0);
      arg_better(best_v, best_i, gi, lo + i);
    }
    __syncthreads();
  }
  for (int i = tid; i < nl; i += MVC_THREADS) u_out[lo + i] = u[i];
  if (tid == 0 && rank == 0 && iters_out) *iters_out = it;
  // peers may still be reading this CTA's candidate / partial buffers
  cluster.sync();
}

}  // namespace nb200

using namespace nb200;

extern "C" {

size_t nb200_mvee_workspace_bytes(int64_t n) {
  return sizeof(double) * (size_t)(n < 1 ? 1 : n);
}

int nb200_mvee_weights(const double* qT_d, int64_t n, int d, int max_updates,
                       double tol, double* u_d, int32_t* iters_d,
                       void* workspace_d, size_t workspace_bytes,
                       void* stream) {
  NB_CHECK(n > d && d >= 1 && d <= NB200_D_MAX,
           "need more points than dimensions, 1 <= d <= 128");
  NB_CHECK(n < (1ll << 31), "too many points");
  NB_CHECK(max_updates >= 0 && tol > 0.0, "bad iteration controls");
  NB_CHECK(workspace_bytes >= nb200_mvee_workspace_bytes(n),
           "workspace too small");
  const int D = d + 1;
  // smallest cluster whose CTAs can keep their slice of the points resident
  for (int C = 1; C <= MVC_MAX_CLUSTER; C *= 2) {
    if (C > n) break;
    const int ld = (int)(((n + C - 1) / C + 3) / 4 * 4);
    const size_t smc = sizeof(double) * (2 * (size_t)D * (D | 1) +
                                         (size_t)(1 + 2 * MVC_MAX_CLUSTER) * D +
                                         (size_t)(d + 2) * ld);
    if (smc > 200 * 1024) continue;
    NB_CUDA(cudaFuncSetAttribute(k_mvee_cluster,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smc));
    NB_CUDA(cudaFuncSetAttribute(
        k_mvee_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed,
        C > 8 ? 1 : 0));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(C, 1, 1);
    cfg.blockDim = dim3(MVC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smc;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = C; attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    if (C > 8) {
      // a 16-CTA cluster is not guaranteed to be schedulable: ask first
      int max_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, k_mvee_cluster,
                                         &cfg) != cudaSuccess ||
          max_clusters < 1) {
        cudaGetLastError();
        break;
      }
    }
    ProfScope prof(ST_FIT, (cudaStream_t)stream);
    NB_CUDA(cudaLaunchKernelEx(&cfg, k_mvee_cluster, qT_d, (int)n, d, ld,
                               max_updates, tol, u_d, iters_d));
    NB_LAUNCH_OK();
    return 0;
  }
  // too many points for a cluster: the single-CTA kernel that streams the
  // points from L2
  const size_t smem = sizeof(double) * ((size_t)D * (D | 1) + 2 * (size_t)D);
  NB_CUDA(cudaFuncSetAttribute(k_mvee,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  ProfScope prof(ST_FIT, (cudaStream_t)stream);
  k_mvee<<<1, MV_THREADS, smem, (cudaStream_t)stream>>>(
      qT_d, n, d, max_updates, tol, u_d, (double*)workspace_d, iters_d);
  NB_LAUNCH_OK();
  return 0;
}

}  // extern "C"
