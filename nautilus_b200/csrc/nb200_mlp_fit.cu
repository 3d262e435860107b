// On-chip trainer of the emulator ensemble.
//
// Replaces NeuralNetworkEmulator.train (nautilus/neural.py:50-98), i.e.
// n_networks x sklearn MLPRegressor.fit with the reference's defaults
// (nautilus/neural.py:79-81): Adam (beta 0.9/0.999, eps 1e-8, bias-corrected
// step; sklearn/neural_network/_stochastic_optimizers.py:255-287), squared
// loss 0.5*mean((y-t)^2) (sklearn/neural_network/_base.py:168-189), ReLU
// hidden layers, identity output, minibatches of min(200, M) rows reshuffled
// every epoch, Glorot-uniform initialisation of weights AND biases with bound
// sqrt(6/(fan_in+fan_out)) (_multilayer_perceptron.py:445-460), stop after
// more than n_iter_no_change epochs without improving the best epoch loss by
// tol (:823-828), at most max_iter epochs.
//
// The reference's fit is ~800 strictly sequential 11-MFLOP Adam steps per
// network -- latency, not FLOPs -- so a host-launched
// step loop would be pure launch overhead.  Here ONE persistent CTA per
// network keeps the weights, the gradient and one chunk of activations in
// shared memory for the whole fit; Adam moments live in L2-resident global
// memory; the epoch loop, the loss curve and the stopping rule run on the
// device.  Arithmetic is fp32 (the reference trains in fp64; parity of fit is
// statistical, tests/test_neural.py:15 of the reference).  Differences that
// are deliberate: Philox instead of MT19937 for init and shuffling, and the
// per-epoch permutation is an affine map i -> (a*i + b) mod M with gcd(a, M)
// = 1 instead of a Fisher-Yates shuffle.
#include <cooperative_groups.h>

#include "nb200_common.cuh"
#include "nb200_rng.cuh"

namespace cg = cooperative_groups;

namespace nb200 {

constexpr int FIT_THREADS = 512;
constexpr int FIT_ROWS_MAX = 64;    // rows of a minibatch resident at once
constexpr int FIT_MAX_LAYERS = 6;   // weight matrices
// Every network is trained by a thread-block CLUSTER of FIT_CLUSTER CTAs: the
// minibatch is split over the CTAs (data parallel), each CTA holds the full
// parameter set, and the minibatch gradient is all-reduced through
// distributed shared memory, in rank order, so that all CTAs apply the same
// Adam step to bit-identical weights.
#ifndef NB200_FIT_CLUSTER
#define NB200_FIT_CLUSTER 8
#endif
constexpr int FIT_CLUSTER = NB200_FIT_CLUSTER;

struct FitArgs {
  int n_lay, d, batch, max_epochs, patience;
  int sizes[FIT_MAX_LAYERS + 1];
  int w_off[FIT_MAX_LAYERS], b_off[FIT_MAX_LAYERS];   // float offsets in W
  int a_off[FIT_MAX_LAYERS + 1], a_stride[FIT_MAX_LAYERS + 1];
  int n_params, delta_off, delta_stride, smem_floats, gsum_off;
  // small mode: transposed copies W^T[n][k] of the layers l >= 1 (row stride
  // wt_ld[l], a multiple of 8) so that the back-propagation of delta reads
  // its eight weights per step as two 16-byte loads
  int wt_base, wt_off[FIT_MAX_LAYERS], wt_ld[FIT_MAX_LAYERS];
  int rows;      // resident chunk (64, 32 or 16: what fits shared memory)
  int big;       // weights / gradient too large for shared memory: they live
                 // in (L2-resident) global memory, activations stay on chip
  int p_quarter; // ceil(n_params / FIT_CLUSTER)
  float lr, beta1, beta2, eps, tol;
  unsigned long long seed;
  long long m;
};

__device__ __forceinline__ long long gcd_ll(long long a, long long b) {
  while (b) { const long long t = a % b; a = b; b = t; }
  return a;
}

// BIG is a template parameter so that in the usual (small) mode every
// parameter / gradient / activation access is provably a shared-memory access
// (LDS/STS instead of generic loads)
template <bool BIG>
__global__ void __cluster_dims__(FIT_CLUSTER, 1, 1)
__launch_bounds__(FIT_THREADS, 1)
k_mlp_fit(const FitArgs A, const float* __restrict__ x,
          const float* __restrict__ y, float* __restrict__ moments,
          double* __restrict__ weights_out, int* __restrict__ n_iter_out,
          double* __restrict__ loss_out, float* __restrict__ big_store) {
  extern __shared__ float fs[];
  // per-CTA parameter storage: shared memory, or global memory in big mode
  // ([W | G | Gq] per CTA, Gq padded to n_params for simple indexing)
  float* big_base = big_store + (size_t)blockIdx.x * 3 * A.n_params;
  float* W = BIG ? big_base : fs;    // parameters of this network
  float* G = BIG ? big_base + A.n_params : fs + A.n_params;
  float* act = BIG ? fs : fs + 2 * A.n_params;   // activations a_0..a_L
  float* dl0 = fs + A.delta_off;       // delta ping-pong
  float* dl1 = dl0 + A.rows * A.delta_stride;
  // this CTA's reduced quarter of G
  float* Gq = BIG ? big_base + 2 * A.n_params : fs + A.gsum_off;
  float* WT = fs + A.wt_base;          // small mode only
  __shared__ float red[FIT_THREADS / 32];
  __shared__ float s_loss;
  __shared__ float s_bsq;              // this CTA's share of sum (y - t)^2

  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int net = blockIdx.x / FIT_CLUSTER;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = FIT_THREADS / 32;
  const int P = A.n_params;
  float* mom_m = moments + (size_t)blockIdx.x * 2 * P;
  float* mom_v = mom_m + P;

  // ---- Glorot-uniform init (weights and biases), zero moments -----------
  for (int l = 0; l < A.n_lay; ++l) {
    const int fi = A.sizes[l], fo = A.sizes[l + 1];
    const float bound = sqrtf(6.0f / (float)(fi + fo));
    const int cnt = fi * fo + fo;
    for (int e = tid; e < cnt; e += FIT_THREADS) {
      const Philox rng((unsigned long long)(A.w_off[l] + e), 0x1000u + net,
                       A.seed);
      const uint4 w = rng.block(0);
      const float u = (float)u01_32(w.x);
      const float v = (2.0f * u - 1.0f) * bound;
      if (e < fi * fo) W[A.w_off[l] + e] = v;
      else W[A.b_off[l] + (e - fi * fo)] = v;
    }
  }
  for (int e = tid; e < P; e += FIT_THREADS) {
    G[e] = 0.f; mom_m[e] = 0.f; mom_v[e] = 0.f;
  }
  __syncthreads();
  if (!BIG) {
    for (int l = 1; l < A.n_lay; ++l) {
      const int fi = A.sizes[l], fo = A.sizes[l + 1], ldt = A.wt_ld[l];
      for (int e = tid; e < fo * ldt; e += FIT_THREADS) {
        const int n = e / ldt, k = e - n * ldt;
        WT[A.wt_off[l] + e] = k < fi ? W[A.w_off[l] + k * fo + n] : 0.f;
      }
    }
    __syncthreads();
  }

  const long long M = A.m;
  const int n_batches = (int)((M + A.batch - 1) / A.batch);
  float best_loss = INFINITY;
  int no_improve = 0, epoch = 0;
  long long t_adam = 0;
  float last_loss = 0.f;

  for (epoch = 0; epoch < A.max_epochs; ++epoch) {
    // affine permutation of the rows for this epoch
    long long pa, pb;
    {
      const Philox rng((unsigned long long)epoch, 0x2000u + net, A.seed);
      const uint4 w = rng.block(0);
      pa = (long long)((((unsigned long long)w.x << 32) | w.y) % (unsigned long long)M);
      pb = (long long)((((unsigned long long)w.z << 32) | w.w) % (unsigned long long)M);
      if (pa == 0) pa = 1;
      while (gcd_ll(pa, M) != 1) pa = pa % M + 1;
    }
    float epoch_loss = 0.f;
    for (int bi = 0; bi < n_batches; ++bi) {
      const long long b_lo = (long long)bi * A.batch;
      const int bn = (int)min((long long)A.batch, M - b_lo);
      float batch_sq = 0.f;            // sum (y - t)^2 over this CTA's rows
      // this CTA's slice of the minibatch
      const int per = (bn + FIT_CLUSTER - 1) / FIT_CLUSTER;
      const int my_lo = min(bn, crank * per), my_hi = min(bn, my_lo + per);
      for (int c_lo = my_lo; c_lo < my_hi; c_lo += A.rows) {
        const int R = min(A.rows, my_hi - c_lo);
        // ---- gather chunk rows: a_0 = x[perm], targets in dl1 tail --------
        float* a0 = act + A.a_off[0];
        for (int e = tid; e < R * A.d; e += FIT_THREADS) {
          const int rr = e / A.d, k = e - rr * A.d;
          const long long src = (pa * (b_lo + c_lo + rr) + pb) % M;
          a0[rr * A.a_stride[0] + k] = x[src * A.d + k];
        }
        __syncthreads();
        // ---- forward ------------------------------------------------------
        for (int l = 0; l < A.n_lay; ++l) {
          const int fi = A.sizes[l], fo = A.sizes[l + 1];
          const float* in = act + A.a_off[l];
          float* out = act + A.a_off[l + 1];
          const int si = A.a_stride[l], so = A.a_stride[l + 1];
          const float* Wl = W + A.w_off[l];
          const float* bl = W + A.b_off[l];
          const int coltiles = (fo + 7) / 8, rowgroups = (R + 31) / 32;
          const bool last = (l == A.n_lay - 1);
          for (int task = warp; task < rowgroups * coltiles; task += nwarps) {
            const int rg = task / coltiles, n0 = (task - rg * coltiles) * 8;
            const int mrow = rg * 32 + lane;
            const bool ok = mrow < R;
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              acc[j] = (n0 + j < fo) ? bl[n0 + j] : 0.f;
            const float* ap = in + mrow * si;
            if (n0 + 8 <= fo) {
              // full tile: no per-element predicates; 8-byte weight loads
              // when the rows of W are 8-byte aligned
              if (!BIG && ((fo | A.w_off[l]) & 1) == 0) {
                const float2* wr2 =
                    reinterpret_cast<const float2*>(Wl + n0);
                const int ld2 = fo >> 1;
#pragma unroll 4
                for (int k = 0; k < fi; ++k) {
                  const float a = ok ? ap[k] : 0.f;
                  const float2 w0 = wr2[k * ld2], w1 = wr2[k * ld2 + 1];
                  const float2 w2 = wr2[k * ld2 + 2], w3 = wr2[k * ld2 + 3];
                  acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]);
                  acc[2] = fmaf(a, w1.x, acc[2]); acc[3] = fmaf(a, w1.y, acc[3]);
                  acc[4] = fmaf(a, w2.x, acc[4]); acc[5] = fmaf(a, w2.y, acc[5]);
                  acc[6] = fmaf(a, w3.x, acc[6]); acc[7] = fmaf(a, w3.y, acc[7]);
                }
              } else {
#pragma unroll 4
                for (int k = 0; k < fi; ++k) {
                  const float a = ok ? ap[k] : 0.f;
                  const float* wr = Wl + k * fo + n0;
#pragma unroll
                  for (int j = 0; j < 8; ++j) acc[j] = fmaf(a, wr[j], acc[j]);
                }
              }
            } else {
              for (int k = 0; k < fi; ++k) {
                const float a = ok ? ap[k] : 0.f;
                const float* wr = Wl + k * fo + n0;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (n0 + j < fo) acc[j] = fmaf(a, wr[j], acc[j]);
              }
            }
            if (ok) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (n0 + j < fo)
                  out[mrow * so + n0 + j] = last ? acc[j] : fmaxf(acc[j], 0.f);
            }
          }
          __syncthreads();
        }
        // ---- output delta and loss ------------------------------------------
        float* dcur = dl0;
        float* dnext = dl1;
        {
          const float* yp = act + A.a_off[A.n_lay];
          const int so = A.a_stride[A.n_lay];
          float sq = 0.f;
          for (int rr = tid; rr < R; rr += FIT_THREADS) {
            const long long src = (pa * (b_lo + c_lo + rr) + pb) % M;
            const float diff = yp[rr * so] - y[src];
            sq += diff * diff;
            dcur[rr * A.delta_stride] = diff / (float)bn;
          }
          sq = warp_sum(sq);
          if (lane == 0) red[warp] = sq;
          __syncthreads();
          if (tid == 0) {
            float t = 0.f;
            for (int w2 = 0; w2 < nwarps; ++w2) t += red[w2];
            s_loss = t;
          }
          __syncthreads();
          batch_sq += s_loss;
        }
        // ---- backward -------------------------------------------------------
        for (int l = A.n_lay - 1; l >= 0; --l) {
          const int fi = A.sizes[l], fo = A.sizes[l + 1];
          const float* ain = act + A.a_off[l];
          const int si = A.a_stride[l];
          float* Gw = G + A.w_off[l];
          float* Gb = G + A.b_off[l];
          // G_w[k][n] += sum_m a[m][k] * delta[m][n]; lanes over n, 8 k's
          {
            const int ktiles = (fi + 7) / 8, ngroups = (fo + 31) / 32;
            for (int task = warp; task < ktiles * ngroups; task += nwarps) {
              const int kt = task / ngroups, ng = task - kt * ngroups;
              const int n = ng * 32 + lane, k0 = kt * 8;
              const bool ok = n < fo;
              float acc[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = 0.f;
              const float* dp = dcur + n;
              if (k0 + 8 <= fi) {
#pragma unroll 4
                for (int mm = 0; mm < R; ++mm) {
                  const float dv = ok ? dp[mm * A.delta_stride] : 0.f;
                  const float* ar = ain + mm * si + k0;
#pragma unroll
                  for (int j = 0; j < 8; ++j) acc[j] = fmaf(ar[j], dv, acc[j]);
                }
              } else {
                for (int mm = 0; mm < R; ++mm) {
                  const float dv = ok ? dp[mm * A.delta_stride] : 0.f;
                  const float* ar = ain + mm * si + k0;
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    if (k0 + j < fi) acc[j] = fmaf(ar[j], dv, acc[j]);
                }
              }
              if (ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (k0 + j < fi) Gw[(k0 + j) * fo + n] += acc[j];
              }
            }
            for (int n = tid; n < fo; n += FIT_THREADS) {
              float s = 0.f;
              for (int mm = 0; mm < R; ++mm) s += dcur[mm * A.delta_stride + n];
              Gb[n] += s;
            }
          }
          // delta_prev[m][k] = (sum_n delta[m][n] W[k][n]) * (a[m][k] > 0)
          if (l > 0) {
            const float* Wl = W + A.w_off[l];
            const int ktiles = (fi + 7) / 8, rowgroups = (R + 31) / 32;
            for (int task = warp; task < rowgroups * ktiles; task += nwarps) {
              const int rg = task / ktiles, k0 = (task - rg * ktiles) * 8;
              const int mrow = rg * 32 + lane;
              const bool ok = mrow < R;
              float acc[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = 0.f;
              const float* dp = dcur + mrow * A.delta_stride;
              const float* wk = Wl + k0 * fo;
              if (!BIG) {
                // W^T[n][k0 .. k0+7]: two 16-byte loads (rows are padded with
                // zeros, so the last tile needs no predicates)
                const float4* wt4 = reinterpret_cast<const float4*>(
                    WT + A.wt_off[l] + k0);
                const int ld4 = A.wt_ld[l] >> 2;
#pragma unroll 4
                for (int n = 0; n < fo; ++n) {
                  const float dv = ok ? dp[n] : 0.f;
                  const float4 wa = wt4[n * ld4], wb = wt4[n * ld4 + 1];
                  acc[0] = fmaf(dv, wa.x, acc[0]); acc[1] = fmaf(dv, wa.y, acc[1]);
                  acc[2] = fmaf(dv, wa.z, acc[2]); acc[3] = fmaf(dv, wa.w, acc[3]);
                  acc[4] = fmaf(dv, wb.x, acc[4]); acc[5] = fmaf(dv, wb.y, acc[5]);
                  acc[6] = fmaf(dv, wb.z, acc[6]); acc[7] = fmaf(dv, wb.w, acc[7]);
                }
              } else if (k0 + 8 <= fi) {
#pragma unroll 4
                for (int n = 0; n < fo; ++n) {
                  const float dv = ok ? dp[n] : 0.f;
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    acc[j] = fmaf(dv, wk[j * fo + n], acc[j]);
                }
              } else {
                for (int n = 0; n < fo; ++n) {
                  const float dv = ok ? dp[n] : 0.f;
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    if (k0 + j < fi) acc[j] = fmaf(dv, wk[j * fo + n], acc[j]);
                }
              }
              if (ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (k0 + j < fi)
                    dnext[mrow * A.delta_stride + k0 + j] =
                        ain[mrow * si + k0 + j] > 0.f ? acc[j] : 0.f;
              }
            }
          }
          __syncthreads();
          float* tmp = dcur; dcur = dnext; dnext = tmp;
        }
      }
      // ---- all-reduce of the gradient and the loss over the cluster -----------
      if (tid == 0) s_bsq = batch_sq;
      if (BIG) __threadfence();      // G lives in global memory
      cluster.sync();
      // reduce-scatter: this CTA sums ITS quarter of the gradient over all
      // ranks (rank order => identical on every CTA) ...
      {
        const int lo = crank * A.p_quarter;
        const int hi = min(P, lo + A.p_quarter);
        for (int e = lo + tid; e < hi; e += FIT_THREADS) {
          float g = 0.f;
#pragma unroll
          for (int q = 0; q < FIT_CLUSTER; ++q) {
            const float* Gp =
                BIG ? big_store + ((size_t)(blockIdx.x - crank + q) * 3 + 1) *
                                      A.n_params
                    : cluster.map_shared_rank(G, q);
            g += BIG ? __ldcg(Gp + e) : Gp[e];
          }
          Gq[e - lo] = g;
        }
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < FIT_CLUSTER; ++q)
          t += *cluster.map_shared_rank(&s_bsq, q);
        batch_sq = t;
      }
      if (BIG) __threadfence();
      cluster.sync();      // ... all-gather happens inside the Adam loop
      // ---- Adam step on the whole minibatch gradient --------------------------
      t_adam += 1;
      const float b1t = powf(A.beta1, (float)t_adam);
      const float b2t = powf(A.beta2, (float)t_adam);
      const float lr_t = A.lr * sqrtf(1.f - b2t) / (1.f - b1t);
      // six parameters per thread and pass: all their loads (the reduced
      // gradient from its owner's shared memory, the two moments from L2) are
      // issued before the first dependent instruction -- the loop used to pay
      // one remote round trip per parameter
      constexpr int AB = 6;
      for (int base = tid; base < P; base += AB * FIT_THREADS) {
        float gqv[AB], mv[AB], vv[AB], wv[AB];
#pragma unroll
        for (int u = 0; u < AB; ++u) {
          const int e = base + u * FIT_THREADS;
          if (e < P) {
            const int owner = e / A.p_quarter;
            gqv[u] =
                BIG ? __ldcg(big_store +
                             ((size_t)(blockIdx.x - crank + owner) * 3 + 2) *
                                 A.n_params + (e - owner * A.p_quarter))
                    : cluster.map_shared_rank(Gq, owner)[e - owner * A.p_quarter];
            mv[u] = mom_m[e];
            vv[u] = mom_v[e];
            wv[u] = W[e];
          }
        }
#pragma unroll
        for (int u = 0; u < AB; ++u) {
          const int e = base + u * FIT_THREADS;
          if (e < P) {
            const float gq = gqv[u];
            const float mq = A.beta1 * mv[u] + (1.f - A.beta1) * gq;
            const float vq = A.beta2 * vv[u] + (1.f - A.beta2) * gq * gq;
            mom_m[e] = mq;
            mom_v[e] = vq;
            const float w_new = wv[u] - lr_t * mq / (sqrtf(vq) + A.eps);
            W[e] = w_new;
            if (!BIG && e >= A.w_off[1]) {
              // keep the transposed copy in step (biases have none)
              int l = 1;
              while (l + 1 < A.n_lay && e >= A.w_off[l + 1]) ++l;
              const int r = e - A.w_off[l], fo = A.sizes[l + 1];
              if (r < A.sizes[l] * fo) {
                const int k = r / fo, n = r - k * fo;
                WT[A.wt_off[l] + n * A.wt_ld[l] + k] = w_new;
              }
            }
            G[e] = 0.f;
          }
        }
      }
      __syncthreads();
      epoch_loss += 0.5f * batch_sq;   // = batch_loss * bn
    }
    last_loss = epoch_loss / (float)M;
    // stopping rule (every thread evaluates the same scalars)
    if (last_loss > best_loss - A.tol) no_improve += 1; else no_improve = 0;
    if (last_loss < best_loss) best_loss = last_loss;
    if (no_improve > A.patience) { epoch += 1; break; }
  }
  // ---- export -------------------------------------------------------------------
  // peers may still be reading this CTA's Gq (distributed shared memory)
  // in their last Adam step: nobody leaves before everybody is done
  cluster.sync();
  double* wo = weights_out + (size_t)net * P;
  if (crank == 0)
    for (int e = tid; e < P; e += FIT_THREADS) wo[e] = (double)W[e];
  if (tid == 0 && crank == 0) {
    n_iter_out[net] = epoch;
    loss_out[net] = (double)last_loss;
  }
}

__global__ void k_f64_to_f32(const double* __restrict__ in, long long n,
                             float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

}  // namespace nb200

namespace nb200 {
// the tcgen05 trainer (nb200_mlp_fit_tc.cu)
struct FitTcArgs;
bool fit_tc_plan(const int32_t* sizes, int n_lay, int64_t m, int batch,
                 FitTcArgs* out);
int launch_fit_tc_from(const int32_t* sizes, int n_lay, int64_t m, int batch,
                       int max_epochs, int patience, float lr, float beta1,
                       float beta2, float eps, float tol,
                       unsigned long long seed, const float* x32,
                       const float* y32, float* moments, int n_net,
                       double* weights_out, int* n_iter_out, double* loss_out,
                       cudaStream_t st, bool* used);
}  // namespace nb200

using namespace nb200;

extern "C" {

size_t nb200_mlp_fit_workspace_bytes(int64_t m, int d, int n_params,
                                     int n_net) {
  // x32, y32, Adam moments (2P) and big-mode parameter storage (3P) per CTA
  return (size_t)m * d * 4 + (size_t)m * 4 +
         (size_t)n_net * FIT_CLUSTER * 5 * n_params * 4 + 1024;
}

int nb200_mlp_fit(const double* x_d, const double* y_d, int64_t m, int d,
                  const int32_t* sizes_h, int n_lay, int n_net, uint64_t seed,
                  double lr, double beta1, double beta2, double eps,
                  int batch_size, int max_epochs, double tol, int patience,
                  double* weights_out_d, int32_t* n_iter_out_d,
                  double* loss_out_d, void* workspace_d,
                  size_t workspace_bytes, void* stream) {
  NB_CHECK(n_lay >= 1 && n_lay <= FIT_MAX_LAYERS, "unsupported layer count");
  NB_CHECK(m >= 1 && d >= 1 && n_net >= 1, "empty training problem");
  NB_CHECK(sizes_h[0] == d && sizes_h[n_lay] == 1, "layer sizes must map d->1");
  cudaStream_t st = (cudaStream_t)stream;
  FitArgs A;
  memset(&A, 0, sizeof(A));
  A.n_lay = n_lay; A.d = d; A.m = m;
  A.batch = (int)(batch_size < m ? batch_size : m);
  A.max_epochs = max_epochs; A.patience = patience;
  A.lr = (float)lr; A.beta1 = (float)beta1; A.beta2 = (float)beta2;
  A.eps = (float)eps; A.tol = (float)tol; A.seed = seed;
  int off = 0, maxw = 1;
  for (int l = 0; l <= n_lay; ++l) {
    A.sizes[l] = sizes_h[l];
    NB_CHECK(sizes_h[l] >= 1 && sizes_h[l] <= NB200_W_MAX, "layer width");
    if (sizes_h[l] > maxw) maxw = sizes_h[l];
  }
  for (int l = 0; l < n_lay; ++l) {
    A.w_off[l] = off; off += A.sizes[l] * A.sizes[l + 1];
    A.b_off[l] = off; off += A.sizes[l + 1];
  }
  A.n_params = off;
  A.delta_stride = maxw | 1;
  A.p_quarter = (A.n_params + FIT_CLUSTER - 1) / FIT_CLUSTER;
  // largest resident chunk that fits: weights + gradient + a quarter-size
  // reduction buffer + activations and two delta buffers of `rows` rows
  for (A.rows = FIT_ROWS_MAX; A.rows >= 8; A.rows >>= 1) {
    int aoff = 0;
    for (int l = 0; l <= n_lay; ++l) {
      A.a_stride[l] = A.sizes[l] | 1;
      A.a_off[l] = aoff;
      aoff += A.rows * A.a_stride[l];
    }
    A.delta_off = 2 * A.n_params + aoff;
    A.gsum_off = A.delta_off + 2 * A.rows * A.delta_stride;
    A.wt_base = (A.gsum_off + A.p_quarter + 3) / 4 * 4;
    int wt = 0;
    for (int l = 1; l < n_lay; ++l) {
      A.wt_ld[l] = (A.sizes[l] + 7) / 8 * 8;
      A.wt_off[l] = wt;
      wt += A.sizes[l + 1] * A.wt_ld[l];
    }
    A.smem_floats = A.wt_base + wt;
    if ((size_t)A.smem_floats * 4 <= 220 * 1024) break;
  }
  if ((size_t)A.smem_floats * 4 > 220 * 1024) {
    // big mode: weights, gradient and the reduction buffer in global memory
    // (a few hundred KB, L2 resident); activations of 64 rows on chip
    A.big = 1;
    for (A.rows = FIT_ROWS_MAX; A.rows >= 8; A.rows >>= 1) {
      int aoff = 0;
      for (int l = 0; l <= n_lay; ++l) {
        A.a_stride[l] = A.sizes[l] | 1;
        A.a_off[l] = aoff;
        aoff += A.rows * A.a_stride[l];
      }
      A.delta_off = aoff;
      A.gsum_off = 0;
      A.smem_floats = A.delta_off + 2 * A.rows * A.delta_stride;
      if ((size_t)A.smem_floats * 4 <= 220 * 1024) break;
    }
  }
  const size_t smem = (size_t)A.smem_floats * 4;
  NB_CHECK(smem <= 220 * 1024,
           "network too wide for the trainer (an 8-row chunk of activations "
           "must fit 220 KB of shared memory)");
  NB_CHECK(workspace_bytes >=
               nb200_mlp_fit_workspace_bytes(m, d, A.n_params, n_net),
           "workspace too small");
  float* x32 = (float*)workspace_d;
  float* y32 = x32 + (size_t)m * d;
  float* moments = y32 + ((m + 3) / 4) * 4;
  float* big_store = moments + (size_t)n_net * FIT_CLUSTER * 2 * A.n_params;
  ProfScope prof(ST_FIT, st);
  k_f64_to_f32<<<(unsigned)((m * d + 255) / 256), 256, 0, st>>>(x_d, m * d, x32);
  NB_LAUNCH_OK();
  k_f64_to_f32<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(y_d, m, y32);
  NB_LAUNCH_OK();
  {
    // tensor-core trainer whenever the problem fits its envelope
    bool used = false;
    const int rc = launch_fit_tc_from(
        sizes_h, n_lay, m, A.batch, max_epochs, patience, A.lr, A.beta1,
        A.beta2, A.eps, A.tol, seed, x32, y32, moments, n_net, weights_out_d,
        n_iter_out_d, loss_out_d, st, &used);
    if (rc) return rc;
    if (used) return 0;
  }
  auto kern = A.big ? k_mlp_fit<true> : k_mlp_fit<false>;
  NB_CUDA(cudaFuncSetAttribute(kern,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  kern<<<n_net * FIT_CLUSTER, FIT_THREADS, smem, st>>>(A, x32, y32, moments,
                                                       weights_out_d,
                                                       n_iter_out_d,
                                                       loss_out_d, big_store);
  NB_LAUNCH_OK();
  return 0;
}

}  // extern "C"
