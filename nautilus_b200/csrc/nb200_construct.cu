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
#include "nb200_common.cuh"

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
