// The emulator trainer on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same algorithm, same random numbers and same cluster layout as k_mlp_fit
// (nb200_mlp_fit.cu: scikit-learn's MLPRegressor.fit with the reference's
// defaults, nautilus/neural.py:50-98 -- Adam, squared loss, ReLU, minibatches,
// patience stopping; one thread-block cluster per network, the minibatch split
// over its CTAs, gradient reduce-scatter / all-gather through distributed
// shared memory), but every matrix product of a step runs as tcgen05.mma
// kind::tf32 with fp32 accumulation in TMEM:
//
//   forward   z_l = a_l W_l + b_l      D[rows x fo]  = A(TMEM: a_l | 1) . WF_l
//   deltas    e_l = (e_l+1 W_l+1^T)    D[rows x fi]  = A(TMEM: e_l+1)   . WB_l+1
//   gradients gW_l = a_l^T e_l         D[fi+1 x fo]  = AT_l(smem) . ET_l(smem)
//
// (the bias is the weight row of a constant-one input column, so its gradient
// is row fi of gW_l).  Activations and deltas live in TMEM with one row of
// the minibatch per lane, exactly like the predict kernel; the epilogue warps
// also leave TRANSPOSED copies a_l^T, e_l^T in shared memory (K-major with the
// minibatch row as K; a 144-byte stride between the 16-byte K chunks keeps
// the 32 lanes of a warp on 32 different banks), which the gradient products
// read as shared-memory operands.  The fp32 master weights ARE the forward
// operand WF_l (kind::tf32 reads the top 19 bits of an fp32 word); Adam updates
// them, and their K-major transposes WB_l, in place.  The fan-out-1 output
// layer and its delta are a few FMAs per row on the CUDA cores.
//
// Envelope (else nb200_mlp_fit falls back to k_mlp_fit): 1-3 hidden layers,
// fan_in + 1 <= 128, hidden widths <= 240, <= 32 minibatch rows per CTA
// (batch <= 256), all operands in 200 KB of shared memory, <= 512 TMEM columns.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "nb200_common.cuh"
#include "nb200_rng.cuh"
#include "nb200_tc.cuh"

namespace cg = cooperative_groups;

namespace nb200 {

constexpr int FT_THREADS = 512;
constexpr int FT_CLUSTER = 8;
constexpr int FT_MAX_HID = 3;
constexpr int FT_ROWS = 32;              // minibatch rows per CTA = K of gW
constexpr int FT_LBO = 144;              // bytes between K chunks (transposes)
constexpr int FT_SBO = 8 * FT_LBO;       // bytes between 8-row groups

struct FitTcArgs {
  int H, d, batch, max_epochs, patience, n_params, p_quarter;
  int fi[FT_MAX_HID], fo[FT_MAX_HID];    // hidden layer l: fi -> fo
  int KP[FT_MAX_HID];                    // round8(fi + 1): forward K
  int NP[FT_MAX_HID];                    // round16(fo): forward / gradient N
  int KB[FT_MAX_HID];                    // round8(fo): K of the delta product
  int NB[FT_MAX_HID];                    // round16(fi): N of the delta product
  int w_off[FT_MAX_HID + 1], b_off[FT_MAX_HID + 1];   // parameter order
  int wf_off[FT_MAX_HID], wb_off[FT_MAX_HID];         // float offsets, smem
  int at_off[FT_MAX_HID], et_off[FT_MAX_HID];         // byte offsets, smem
  int g_off, gq_off, wout_off, misc_off, smem_bytes;  // float offsets
  int a_col[FT_MAX_HID + 1];             // TMEM: a_0, then z_l / a_l+1
  int e_col[FT_MAX_HID];                 // TMEM: e_l
  float lr, beta1, beta2, eps, tol;
  unsigned long long seed;
  long long m;
};

__device__ __forceinline__ long long ft_gcd(long long a, long long b) {
  while (b) { const long long t = a % b; a = b; b = t; }
  return a;
}

// D[tmem] (+)= A[smem desc] . B[smem desc]^T, kind::tf32, M = 128
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc,
                                            uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// round an fp32 value to tf32 (half up; the MMA reads the top 19 bits)
__device__ __forceinline__ float tf32_rn(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}
// element (row, k) of a K-major operand: weights (LBO 128, SBO = kp * 32)
__device__ __forceinline__ int wk_index(int row, int k, int kp) {
  return (row >> 3) * (kp * 8) + (k >> 2) * 32 + (row & 7) * 4 + (k & 3);
}
// ... and of a transposed activation / delta (K = minibatch row)
__device__ __forceinline__ int tr_byte(int row, int r) {
  return (row >> 3) * FT_SBO + (r >> 2) * FT_LBO + (row & 7) * 16 + (r & 3) * 4;
}

__global__ void __cluster_dims__(FT_CLUSTER, 1, 1)
__launch_bounds__(FT_THREADS, 1)
k_mlp_fit_tc(const FitTcArgs A, const float* __restrict__ x,
             const float* __restrict__ y, float* __restrict__ moments,
             double* __restrict__ weights_out, int* __restrict__ n_iter_out,
             double* __restrict__ loss_out) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* fs = reinterpret_cast<float*>(smem);
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bsq;
  __shared__ float s_part[4][FT_ROWS];

  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int net = blockIdx.x / FT_CLUSTER;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = A.H, P = A.n_params;
  float* G = fs + A.g_off;
  float* Gq = fs + A.gq_off;
  float* wout = fs + A.wout_off;         // W_H[fo_last] then b_H
  float* mom_m = moments + (size_t)blockIdx.x * 2 * P;
  float* mom_v = mom_m + P;
  const int fo_last = A.fo[H - 1];

  // parameter e (scikit-learn order: W_0, b_0, W_1, ...) -> where it lives
  auto param_ptr = [&](int e, float** wb_copy) -> float* {
    *wb_copy = nullptr;
    int l = 0;
    while (l < H && e >= A.w_off[l + 1]) ++l;
    if (l == H) {                                  // output layer
      const int r = e - A.w_off[H];
      return wout + r;                             // W_H rows, then b_H
    }
    const int r = e - A.w_off[l], fo = A.fo[l];
    float* WF = fs + A.wf_off[l];
    if (r >= A.fi[l] * fo)                         // bias: the row of the one
      return WF + wk_index(r - A.fi[l] * fo, A.fi[l], A.KP[l]);
    const int i = r / fo, o = r - i * fo;
    if (l > 0) *wb_copy = fs + A.wb_off[l] + wk_index(i, o, A.KB[l]);
    return WF + wk_index(o, i, A.KP[l]);
  };

  // ---- setup: zero the operands, Glorot init, TMEM ---------------------------
  for (int e = tid; e < A.smem_bytes / 4; e += FT_THREADS) fs[e] = 0.f;
  if (tid == 0) {
    mbar_init(&mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  for (int l = 0; l <= H; ++l) {
    const int fi = l < H ? A.fi[l] : fo_last, fo = l < H ? A.fo[l] : 1;
    const float bound = sqrtf(6.0f / (float)(fi + fo));
    const int cnt = fi * fo + fo;
    for (int e = tid; e < cnt; e += FT_THREADS) {
      // (the same draws as k_mlp_fit: same initial networks)
      const Philox rng((unsigned long long)(A.w_off[l] + e), 0x1000u + net,
                       A.seed);
      const uint4 w = rng.block(0);
      const float v = (2.0f * (float)u01_32(w.x) - 1.0f) * bound;
      const int pe = e < fi * fo ? A.w_off[l] + e : A.b_off[l] + (e - fi * fo);
      float* wb;
      float* p = param_ptr(pe, &wb);
      *p = v;
      if (wb) *wb = v;
    }
  }
  for (int e = tid; e < P; e += FT_THREADS) { mom_m[e] = 0.f; mom_v[e] = 0.f; }
  if (warp == 0) {
    asm volatile(
        "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
        ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"
                 ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;       // lanes 0..31 hold the rows
  const uint32_t smem_base = smem_u32(smem);
  uint32_t phase = 0;

  // the four warps that own TMEM lanes 0..31 (warp % 4 == 0) run the
  // epilogues, each on a quarter of the columns
  const bool epi = (warp & 3) == 0;
  const int eq = warp >> 2;              // 0..3: which quarter

  // issue helpers (one thread)
  auto issue_ts = [&](int d_col, int a_col, int n, int ksteps, int wf_off,
                      int kp) {
    const uint64_t desc =
        smem_desc(smem_base + 4u * (uint32_t)wf_off, 128u, (uint32_t)kp * 32u);
    const uint32_t id = idesc_tf32(n);
    for (int s = 0; s < ksteps; ++s)
      mma_tf32_ts(tmem + (uint32_t)d_col, tmem + (uint32_t)(a_col + 8 * s),
                  desc + (uint64_t)(16 * s), id, s > 0 ? 1u : 0u);
  };
  auto issue_ss = [&](int d_col, int at_byte, int et_byte, int n) {
    const uint64_t da = smem_desc(smem_base + (uint32_t)at_byte, FT_LBO, FT_SBO);
    const uint64_t db = smem_desc(smem_base + (uint32_t)et_byte, FT_LBO, FT_SBO);
    const uint32_t id = idesc_tf32(n);
    for (int s = 0; s < FT_ROWS / 8; ++s)      // K = 32 rows: 4 steps of 8
      mma_tf32_ss(tmem + (uint32_t)d_col, da + (uint64_t)(s * 2 * FT_LBO / 16),
                  db + (uint64_t)(s * 2 * FT_LBO / 16), id, s > 0 ? 1u : 0u);
  };
  // everything written so far (TMEM by tcgen05.st, shared memory by ordinary
  // stores) is visible to the MMAs issued after this
  auto publish = [&]() {
    tmem_wait_st();
    tc_fence_before();
    fence_async_smem();
    __syncthreads();
  };
  auto await = [&]() {
    mbar_wait(&mbar, phase);
    phase ^= 1u;
    tc_fence_after();
  };

  const long long M = A.m;
  const int n_batches = (int)((M + A.batch - 1) / A.batch);
  float best_loss = INFINITY, last_loss = 0.f;
  int no_improve = 0, epoch = 0;
  long long t_adam = 0;

  for (epoch = 0; epoch < A.max_epochs; ++epoch) {
    long long pa, pb;
    {
      const Philox rng((unsigned long long)epoch, 0x2000u + net, A.seed);
      const uint4 w = rng.block(0);
      pa = (long long)((((unsigned long long)w.x << 32) | w.y) %
                       (unsigned long long)M);
      pb = (long long)((((unsigned long long)w.z << 32) | w.w) %
                       (unsigned long long)M);
      if (pa == 0) pa = 1;
      while (ft_gcd(pa, M) != 1) pa = pa % M + 1;
    }
    float epoch_loss = 0.f;
    for (int bi = 0; bi < n_batches; ++bi) {
      const long long b_lo = (long long)bi * A.batch;
      const int bn = (int)min((long long)A.batch, M - b_lo);
      const int per = (bn + FT_CLUSTER - 1) / FT_CLUSTER;
      const int my_lo = min(bn, crank * per), my_hi = min(bn, my_lo + per);
      const int R = my_hi - my_lo;               // <= FT_ROWS rows, lane = row
      const bool row_ok = lane < R;
      const long long src =
          row_ok ? (pa * (b_lo + my_lo + lane) + pb) % M : 0;

      // ---- a_0 = (x | 1): TMEM (A operand) and its transpose -----------------
      if (epi) {
        const int kp = A.KP[0];
        for (int c = eq * 8; c < kp; c += 32) {
          uint32_t v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int k = c + q;
            float f = 0.f;
            if (row_ok) f = k < A.d ? tf32_rn(x[src * A.d + k])
                                    : (k == A.d ? 1.0f : 0.f);
            v[q] = __float_as_uint(f);
            if (k <= A.d)
              *reinterpret_cast<float*>(smem + A.at_off[0] + tr_byte(k, lane)) =
                  f;
          }
          tmem_st8(tmem + (uint32_t)(A.a_col[0] + c), v);
        }
      }
      publish();
      // ---- forward ---------------------------------------------------------------
      for (int l = 0; l < H; ++l) {
        if (tid == 0) {
          tc_fence_after();
          issue_ts(A.a_col[l + 1], A.a_col[l], A.NP[l], A.KP[l] >> 3,
                   A.wf_off[l], A.KP[l]);
          mma_commit(&mbar);
        }
        await();
        if (epi) {
          // ReLU -> a_l+1 (with its constant-one column) in place, and its
          // transpose; columns beyond the next layer's K are never read
          const int fo = A.fo[l];
          const int width = l + 1 < H ? max(A.NP[l], A.KP[l + 1]) : A.NP[l];
          for (int c = eq * 8; c < width; c += 32) {
            uint32_t v[8];
            if (c < A.NP[l]) {
              tmem_ld8(tmem + (uint32_t)(A.a_col[l + 1] + c), v);
              tmem_wait_ld();
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = 0u;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int k = c + q;
              float f = k < fo ? tf32_rn(fmaxf(__uint_as_float(v[q]), 0.f))
                               : (k == fo ? 1.0f : 0.f);
              if (!row_ok) f = k == fo ? 1.0f : 0.f;
              v[q] = __float_as_uint(f);
              if (l + 1 < H && k <= fo)
                *reinterpret_cast<float*>(smem + A.at_off[l + 1] +
                                          tr_byte(k, lane)) = f;
            }
            tmem_st8(tmem + (uint32_t)(A.a_col[l + 1] + c), v);
          }
        }
        publish();
      }
      // ---- output layer, loss, e_H-1 (CUDA cores; fan_out 1) ---------------------
      // warp 4 q handles columns [8 q, 8 q + 8) + 32 j of a_H; partial dot
      // products meet in shared memory
      float e_out = 0.f;                      // (y - t) / batch of this row
      if (epi) {
        float acc = 0.f;
        for (int c = eq * 8; c < A.NP[H - 1]; c += 32) {
          uint32_t v[8];
          tmem_ld8(tmem + (uint32_t)(A.a_col[H] + c), v);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (c + q < fo_last)
              acc = fmaf(__uint_as_float(v[q]), wout[c + q], acc);
        }
        s_part[eq][lane] = acc;
      }
      __syncthreads();
      float batch_sq = 0.f;
      if (epi) {
        const float yhat = s_part[0][lane] + s_part[1][lane] +
                           s_part[2][lane] + s_part[3][lane] + wout[fo_last];
        const float diff = row_ok ? yhat - y[src] : 0.f;
        e_out = diff / (float)bn;
        if (eq == 0) {
          float sq = warp_sum(diff * diff);
          float gb = warp_sum(e_out);
          if (lane == 0) { batch_sq = sq; G[A.b_off[H]] = gb; }
        }
        // gW_H[i] = sum_r a_H[r][i] e_out[r];  e_H-1 = e_out W_H (a_H > 0)
        const int kb = A.KB[H - 1];
        for (int c = eq * 8; c < max(kb, A.NP[H - 1]); c += 32) {
          uint32_t v[8];
          if (c < A.NP[H - 1]) {
            tmem_ld8(tmem + (uint32_t)(A.a_col[H] + c), v);
            tmem_wait_ld();
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = 0u;
          }
          uint32_t ev[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int k = c + q;
            const float a = k < fo_last ? __uint_as_float(v[q]) : 0.f;
            const float gw = warp_sum(a * e_out);
            if (lane == 0 && k < fo_last) G[A.w_off[H] + k] = gw;
            const float e = (k < fo_last && a > 0.f)
                                ? tf32_rn(e_out * wout[k]) : 0.f;
            ev[q] = __float_as_uint(e);
            if (k < A.NP[H - 1])
              *reinterpret_cast<float*>(smem + A.et_off[H - 1] +
                                        tr_byte(k, lane)) = e;
          }
          if (c < kb) tmem_st8(tmem + (uint32_t)(A.e_col[H - 1] + c), ev);
        }
      }
      publish();
      // ---- deltas of the earlier hidden layers -----------------------------------
      for (int l = H - 2; l >= 0; --l) {
        if (tid == 0) {
          tc_fence_after();
          // e_l = e_l+1 . W_l+1^T : N = round16(fo_l) = NB[l+1], K = KB[l+1]
          issue_ts(A.e_col[l], A.e_col[l + 1], A.NB[l + 1], A.KB[l + 1] >> 3,
                   A.wb_off[l + 1], A.KB[l + 1]);
          mma_commit(&mbar);
        }
        await();
        if (epi) {
          const int fo = A.fo[l];
          for (int c = eq * 8; c < A.NB[l + 1]; c += 32) {
            uint32_t v[8], a[8];
            tmem_ld8(tmem + (uint32_t)(A.e_col[l] + c), v);
            tmem_ld8(tmem + (uint32_t)(A.a_col[l + 1] + c), a);
            tmem_wait_ld();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int k = c + q;
              const float e = (k < fo && row_ok && __uint_as_float(a[q]) > 0.f)
                                  ? tf32_rn(__uint_as_float(v[q])) : 0.f;
              v[q] = __float_as_uint(e);
              *reinterpret_cast<float*>(smem + A.et_off[l] +
                                        tr_byte(k, lane)) = e;
            }
            if (l > 0) tmem_st8(tmem + (uint32_t)(A.e_col[l] + c), v);
          }
        }
        publish();
      }
      // ---- gradients gW_l = (a_l | 1)^T e_l : one MMA group per layer ------------
      if (tid == 0) {
        tc_fence_after();
        for (int l = 0; l < H; ++l)
          issue_ss(A.a_col[l + 1], A.at_off[l], A.et_off[l], A.NP[l]);
        mma_commit(&mbar);
      }
      await();
      {
        // D[lane = input i (fi = the bias row)][col = output o] -> G
        const int qd = warp & 3;            // TMEM lanes 32 qd .. 32 qd + 31
        const int part = warp >> 2;
        const int i = qd * 32 + lane;
        for (int l = 0; l < H; ++l) {
          const int fi = A.fi[l], fo = A.fo[l];
          if (qd * 32 > fi) continue;       // this quadrant holds no row
          for (int c = part * 8; c < A.NP[l]; c += 32) {
            uint32_t v[8];
            tmem_ld8(tmem + ((uint32_t)(qd * 32) << 16) +            (uint32_t)(A.a_col[l + 1] + c), v);
            tmem_wait_ld();
            if (i <= fi) {
              float* dst = i < fi ? G + A.w_off[l] + i * fo : G + A.b_off[l];
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (c + q < fo) dst[c + q] = __uint_as_float(v[q]);
            }
          }
        }
      }
      // ---- all-reduce of the gradient and the loss over the cluster --------------
      if (tid == 0) s_bsq = batch_sq;
      tc_fence_before();
      cluster.sync();
      {
        const int lo = crank * A.p_quarter;
        const int hi = min(P, lo + A.p_quarter);
        for (int e = lo + tid; e < hi; e += FT_THREADS) {
          float g = 0.f;
#pragma unroll
          for (int q = 0; q < FT_CLUSTER; ++q)
            g += cluster.map_shared_rank(G, q)[e];
          Gq[e - lo] = g;
        }
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < FT_CLUSTER; ++q)
          t += *cluster.map_shared_rank(&s_bsq, q);
        batch_sq = t;
      }
      cluster.sync();
      // ---- Adam on the whole minibatch gradient -----------------------------------
      t_adam += 1;
      const float b1t = powf(A.beta1, (float)t_adam);
      const float b2t = powf(A.beta2, (float)t_adam);
      const float lr_t = A.lr * sqrtf(1.f - b2t) / (1.f - b1t);
      constexpr int AB = 6;
      for (int base = tid; base < P; base += AB * FT_THREADS) {
        float gqv[AB], mv[AB], vv[AB];
#pragma unroll
        for (int u = 0; u < AB; ++u) {
          const int e = base + u * FT_THREADS;
          if (e < P) {
            const int owner = e / A.p_quarter;
            gqv[u] = cluster.map_shared_rank(Gq, owner)[e - owner * A.p_quarter];
            mv[u] = mom_m[e];
            vv[u] = mom_v[e];
          }
        }
#pragma unroll
        for (int u = 0; u < AB; ++u) {
          const int e = base + u * FT_THREADS;
          if (e < P) {
            const float gq = gqv[u];
            const float mq = A.beta1 * mv[u] + (1.f - A.beta1) * gq;
            const float vq = A.beta2 * vv[u] + (1.f - A.beta2) * gq * gq;
            mom_m[e] = mq;
            mom_v[e] = vq;
            float* wb;
            float* p = param_ptr(e, &wb);
            const float w_new = *p - lr_t * mq / (sqrtf(vq) + A.eps);
            *p = w_new;
            if (wb) *wb = w_new;
          }
        }
      }
      tc_fence_after();
      __syncthreads();
      epoch_loss += 0.5f * batch_sq;
    }
    last_loss = epoch_loss / (float)M;
    if (last_loss > best_loss - A.tol) no_improve += 1; else no_improve = 0;
    if (last_loss < best_loss) best_loss = last_loss;
    if (no_improve > A.patience) { epoch += 1; break; }
  }
  // ---- export (parameter order, fp64) ---------------------------------------------
  cluster.sync();
  if (crank == 0) {
    double* wo = weights_out + (size_t)net * P;
    for (int e = tid; e < P; e += FT_THREADS) {
      float* wb;
      wo[e] = (double)*param_ptr(e, &wb);
    }
    if (tid == 0) {
      n_iter_out[net] = epoch;
      loss_out[net] = (double)last_loss;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem_slot), "r"(512) : "memory");
}

static inline int r8(int v) { return (v + 7) / 8 * 8; }
static inline int r16(int v) { return (v + 15) / 16 * 16; }
static inline int r32(int v) { return (v + 31) / 32 * 32; }

// Fill the argument block; false if the problem is outside the envelope.
bool fit_tc_plan(const int32_t* sizes, int n_lay, int64_t m, int batch,
                 FitTcArgs* out) {
  const char* env = getenv("NB200_FIT");
  if (env && strcmp(env, "ffma") == 0) return false;
  FitTcArgs A;
  memset(&A, 0, sizeof(A));
  const int H = n_lay - 1;
  if (H < 1 || H > FT_MAX_HID || sizes[n_lay] != 1) return false;
  if ((batch + FT_CLUSTER - 1) / FT_CLUSTER > FT_ROWS) return false;
  A.H = H; A.d = sizes[0]; A.m = m;
  int off = 0;
  for (int l = 0; l <= H; ++l) {
    A.w_off[l] = off; off += sizes[l] * sizes[l + 1];
    A.b_off[l] = off; off += sizes[l + 1];
  }
  A.n_params = off;
  A.p_quarter = (off + FT_CLUSTER - 1) / FT_CLUSTER;
  int fl = 0, col = 0;                       // shared-memory floats, TMEM cols
  for (int l = 0; l < H; ++l) {
    A.fi[l] = sizes[l]; A.fo[l] = sizes[l + 1];
    if (A.fi[l] + 1 > 128 || A.fo[l] > 240) return false;
    A.KP[l] = r8(A.fi[l] + 1); A.NP[l] = r16(A.fo[l]);
    A.KB[l] = r8(A.fo[l]); A.NB[l] = r16(A.fi[l]);
  }
  // transposed operands first: the gradient MMAs read 128 rows (16 groups) of
  // every AT_l; what lies behind the rows that exist is other, readable data
  int bytes = 0;
  for (int l = 0; l < H; ++l) {
    A.at_off[l] = bytes; bytes += (A.fi[l] + 1 + 7) / 8 * FT_SBO;
  }
  for (int l = 0; l < H; ++l) {
    A.et_off[l] = bytes; bytes += A.NP[l] / 8 * FT_SBO;
  }
  bytes = (bytes + 127) / 128 * 128;
  fl = bytes / 4;
  for (int l = 0; l < H; ++l) {
    A.wf_off[l] = fl; fl += A.NP[l] * A.KP[l];
    fl = (fl + 31) / 32 * 32;
  }
  for (int l = 1; l < H; ++l) {
    A.wb_off[l] = fl; fl += A.NB[l] * A.KB[l];
    fl = (fl + 31) / 32 * 32;
  }
  A.g_off = fl; fl += A.n_params;
  A.gq_off = fl; fl += A.p_quarter;
  A.wout_off = fl; fl += A.fo[H - 1] + 1;
  fl = (fl + 31) / 32 * 32;
  A.smem_bytes = fl * 4;
  // (the over-read of AT_l: 16 groups from its base must stay inside)
  if (A.at_off[H - 1] + 16 * FT_SBO > A.smem_bytes) return false;
  if (A.smem_bytes > 200 * 1024) return false;
  A.a_col[0] = col; col += r32(A.KP[0]);
  for (int l = 0; l < H; ++l) {
    A.a_col[l + 1] = col;
    col += r32(l + 1 < H ? (A.NP[l] > A.KP[l + 1] ? A.NP[l] : A.KP[l + 1])
                         : A.NP[l]);
  }
  for (int l = H - 1; l >= 0; --l) {
    A.e_col[l] = col;
    col += r32(l == H - 1 ? (A.KB[l] > 8 ? A.KB[l] : 8) : A.NB[l + 1]);
  }
  if (col > 512) return false;
  *out = A;
  return true;
}

int launch_fit_tc(FitTcArgs A, const float* x32, const float* y32,
                  float* moments, int n_net, double* weights_out,
                  int* n_iter_out, double* loss_out, cudaStream_t st) {
  NB_CUDA(cudaFuncSetAttribute(k_mlp_fit_tc,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               A.smem_bytes));
  k_mlp_fit_tc<<<n_net * FT_CLUSTER, FT_THREADS, A.smem_bytes, st>>>(
      A, x32, y32, moments, weights_out, n_iter_out, loss_out);
  NB_LAUNCH_OK();
  return 0;
}

int launch_fit_tc_from(const int32_t* sizes, int n_lay, int64_t m, int batch,
                       int max_epochs, int patience, float lr, float beta1,
                       float beta2, float eps, float tol,
                       unsigned long long seed, const float* x32,
                       const float* y32, float* moments, int n_net,
                       double* weights_out, int* n_iter_out, double* loss_out,
                       cudaStream_t st, bool* used) {
  FitTcArgs A;
  *used = false;
  if (!fit_tc_plan(sizes, n_lay, m, batch, &A)) return 0;
  A.batch = batch; A.max_epochs = max_epochs; A.patience = patience;
  A.lr = lr; A.beta1 = beta1; A.beta2 = beta2; A.eps = eps; A.tol = tol;
  A.seed = seed;
  *used = true;
  return launch_fit_tc(A, x32, y32, moments, n_net, weights_out, n_iter_out,
                       loss_out, st);
}

}  // namespace nb200
