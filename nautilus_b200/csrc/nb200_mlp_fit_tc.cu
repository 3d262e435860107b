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
//   gradients gW_l^T = e_l^T a_l       D[fo x fi+1]  = ET_l(smem) . AT_l(smem)
//
// (the bias is the weight row of a constant-one input column, so its gradient
// is column fi of gW_l^T).  Activations and deltas live in TMEM, one row of
// the minibatch per lane; the epilogues also leave TRANSPOSED copies a_l^T,
// e_l^T in shared memory (K-major with the minibatch row as K; a 144-byte
// stride between the 16-byte K chunks keeps the lanes of a warp on different
// banks), which the gradient products read as shared-memory operands.
//
// The gradient comes out TRANSPOSED (output unit per lane, inputs along the
// columns), which is the K-major order of the forward operand WF_l: a lane's
// eight consecutive columns are two 16-byte chunks of WF_l, so the gradient
// epilogue writes the gradient straight INTO WF_l with conflict-free 128-bit
// stores (the weights are not lost: every CTA keeps the fp32 master copy of
// the share of the parameters it owns).  The reduce-scatter, Adam and the
// all-gather then run over the WF index space with 128-bit distributed
// shared memory loads / stores: CTA r sums its share of the eight gradients,
// updates its master weights and moments, and stores the new weights into
// the WF of all eight CTAs -- no parameter-order buffer, no scatter.  The
// K-major transposes WB_l of the hidden-to-hidden weights are refreshed by a
// local shared-memory transpose.  The fan-out-1 output layer, its delta and
// its 21 parameters are a few FMAs per row on the CUDA cores, replicated on
// every CTA.
//
// Envelope (else nb200_mlp_fit falls back to k_mlp_fit): 1-3 hidden layers,
// fan_in + 1 <= 128, hidden widths <= 128, <= 32 minibatch rows per CTA
// (batch <= 256), all operands in 200 KB of shared memory, <= 512 TMEM columns.
#include <cooperative_groups.h>
#include <type_traits>
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
constexpr int FT_MAX_W = 128;            // widest hidden layer (a TMEM lane each)
constexpr int FT_LBO = 144;              // bytes between K chunks (transposes)
constexpr int FT_SBO = 8 * FT_LBO;       // bytes between 8-row groups

struct FitTcArgs {
  int H, d, batch, max_epochs, patience, n_params;
  int fi[FT_MAX_HID], fo[FT_MAX_HID];    // hidden layer l: fi -> fo
  int KP[FT_MAX_HID];                    // round8(fi + 1): forward K
  int NP[FT_MAX_HID];                    // round16(fo + 1): forward N (the
                                         // spare row makes the constant one)
  int KB[FT_MAX_HID];                    // round8(fo): K of the delta product
  int NB[FT_MAX_HID];                    // round16(fi): N of the delta product
  int NG[FT_MAX_HID + 1];                // round16(fi + 1): N of the gradient
  int w_off[FT_MAX_HID + 1], b_off[FT_MAX_HID + 1];   // parameter order
  int wf_off[FT_MAX_HID], wb_off[FT_MAX_HID];         // float offsets, smem
  int at_off[FT_MAX_HID + 1], et_off[FT_MAX_HID + 1]; // byte offsets, smem
                                         // (entry H: the output layer)
  // float offsets: Adam state of this CTA's share (m | v | master weights),
  // output layer (weights, gradient, m, v), prefetched minibatch rows
  int own_off, wout_off, gout_off, mout_off, misc_off, scr_off, smem_bytes;
  int wf_floats, share4;                 // WF index space; float4s per CTA
  unsigned int magic[FT_MAX_HID];        // ceil(2^32 / (8 KP)): r / (8 KP)
  int a_col[FT_MAX_HID + 1];             // TMEM: a_0, then z_l / a_l+1
  int e_col[FT_MAX_HID];                 // TMEM: e_l
  int g_col[FT_MAX_HID + 1];             // TMEM: gW_l^T (over the dead a / e)
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
// round an fp32 value to tf32 (nearest; the MMA reads the top 19 bits)
__device__ __forceinline__ float tf32_rn(float v) {
  return __uint_as_float(to_tf32(v));
}
// the same on the bit pattern, for finite non-negative values (two integer
// instructions instead of the guarded conversion)
__device__ __forceinline__ uint32_t tf32_bits(float v) {
  return (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;
}
// element (row, k) of a K-major operand: weights (LBO 128, SBO = kp * 32)
__device__ __forceinline__ int wk_index(int row, int k, int kp) {
  return (row >> 3) * (kp * 8) + (k >> 2) * 32 + (row & 7) * 4 + (k & 3);
}
// ... and of a transposed activation / delta (K = minibatch row)
__device__ __forceinline__ int tr_byte(int row, int r) {
  return (row >> 3) * FT_SBO + (r >> 2) * FT_LBO + (row & 7) * 16 + (r & 3) * 4;
}
// 32-bit shared-memory store at a precomputed shared address: keeps the
// transposed stores of the epilogues at one STS each (written through a
// generic pointer the compiler re-derived the address, cluster CTA id
// included, around every single store)
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)
               : "memory");
  return v;
}
// the same address in CTA `rank` of the cluster, and a 128-bit load from it
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
               : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ldc_v4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void sts_s64(uint32_t addr, long long v) {
  asm volatile("st.shared.s64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ long long lds_s64(uint32_t addr) {
  long long v;
  asm volatile("ld.shared.s64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
// mbar_wait / mma_commit / expect_tx of nb200_tc.cuh on a shared ADDRESS
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (it > (1u << 22)) __trap();   // never hang the GPU on a lost arrival
  }
}
__device__ __forceinline__ void mma_commit_a(uint32_t addr) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 "
      "[%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// H (the number of hidden layers) is a template parameter and every loop over
// the layers is unrolled: the per-layer sizes and offsets are then direct
// constant-bank operands.  (Indexed with a run-time layer they were re-loaded
// by dependent LDCs around every store of the epilogues, ~2000 clocks per
// 8-column chunk.)
template <int H>
__global__ void __cluster_dims__(FT_CLUSTER, 1, 1)
__launch_bounds__(FT_THREADS, 1)
k_mlp_fit_tc(const FitTcArgs A, const float* __restrict__ x,
             const float* __restrict__ y, float* __restrict__ moments,
             double* __restrict__ weights_out, int* __restrict__ n_iter_out,
             double* __restrict__ loss_out) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* fs = reinterpret_cast<float*>(smem);
  // The 32-bit shared address of the dynamic shared memory, PINNED in a
  // register by a volatile asm: in a cluster kernel the conversion reads the
  // CTA's rank (S2UR SR_CgaCtaId, > 100 clocks), and the compiler
  // re-materialised it in front of every single shared-memory access that
  // went through a pointer -- the hot paths below address shared memory as
  // smem_base + offset with explicit ld.shared / st.shared instead.
  uint32_t smem_base;
  asm volatile("{\n\t.reg .u64 t;\n\tcvta.to.shared.u64 t, %1;\n\t"
               "cvt.u32.u64 %0, t;\n\t}"
               : "=r"(smem_base) : "l"(smem));
  __shared__ uint64_t mbar;      // MMA groups
  __shared__ uint64_t mbar_ag;   // all-gather of the new weights (TMA bytes)
  __shared__ uint32_t tmem_slot;
  (void)moments;                         // (the SIMT trainer's global state)

  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int net = blockIdx.x / FT_CLUSTER;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = A.n_params;
  const int fo_last = A.fo[H - 1];
  float* wout = fs + A.wout_off;         // W_H[fo_last] then b_H
  float* gout = fs + A.gout_off;         // their gradient (this CTA's rows)
  float* mout = fs + A.mout_off;         // their Adam moments: m then v
  float* WFall = fs + A.wf_off[0];       // WF_0 | WF_1 | ... (contiguous)
  // this CTA's share of the WF index space: float4s [lo4, hi4)
  const int lo4 = crank * A.share4;
  const int hi4 = min(A.wf_floats / 4, lo4 + A.share4);
  float4* own_m = reinterpret_cast<float4*>(fs + A.own_off);
  float4* own_v = own_m + A.share4;
  float4* own_w = own_v + A.share4;      // fp32 master weights of the share

  // parameter e (scikit-learn order: W_0, b_0, W_1, ...) -> where it lives
  auto param_ptr = [&](int e, float** wb_copy) -> float* {
    *wb_copy = nullptr;
    if (e >= A.w_off[H]) return wout + (e - A.w_off[H]);   // W_H, then b_H
    float* out = nullptr;
#pragma unroll
    for (int l = 0; l < H; ++l) {
      if (e >= A.w_off[l] && e < A.w_off[l + 1]) {
        const int r = e - A.w_off[l], fo = A.fo[l];
        float* WF = fs + A.wf_off[l];
        if (r >= A.fi[l] * fo) {                   // bias: the row of the one
          out = WF + wk_index(r - A.fi[l] * fo, A.fi[l], A.KP[l]);
        } else {
          const int i = r / fo, o = r - i * fo;
          if (l > 0) *wb_copy = fs + A.wb_off[l] + wk_index(i, o, A.KB[l]);
          out = WF + wk_index(o, i, A.KP[l]);
        }
      }
    }
    return out;
  };
  // WB_l (i, o) = WF_l (o, i) for the hidden-to-hidden layers, from this
  // CTA's own WF: 128-bit loads (one output o, inputs i0 .. i0 + 3) scattered
  // into the K-major transpose
  auto refresh_wb = [&]() {
#pragma unroll
    for (int l = 1; l < H; ++l) {
      const int fi = A.fi[l], fo = A.fo[l], kp8 = A.KP[l] * 8;
      for (int t = tid; t < A.NP[l] * A.KP[l] / 4; t += FT_THREADS) {
        // r = 4 t = (o / 8) * (8 KP) + (i / 4) * 32 + (o % 8) * 4
        const int r = 4 * t;
        const int ob = (int)__umulhi((unsigned)r, A.magic[l]);
        const int rem = r - ob * kp8;
        const int o = 8 * ob + ((rem & 31) >> 2), i0 = (rem >> 5) * 4;
        if (o >= fo || i0 >= fi) continue;
        const float4 w =
            lds_v4(smem_base + 4u * (uint32_t)A.wf_off[l] + 16u * (uint32_t)t);
        const uint32_t dst =
            smem_base +
            4u * (uint32_t)(A.wb_off[l] + wk_index(i0, o, A.KB[l]));
        sts_f32(dst, w.x);
        if (i0 + 1 < fi) sts_f32(dst + 16, w.y);
        if (i0 + 2 < fi) sts_f32(dst + 32, w.z);
        if (i0 + 3 < fi) sts_f32(dst + 48, w.w);
      }
    }
  };

  // ---- setup: zero the operands, Glorot init, TMEM ---------------------------
  for (int e = tid; e < A.smem_bytes / 4; e += FT_THREADS) fs[e] = 0.f;
  if (tid == 0) {
    mbar_init(&mbar, 1);
    mbar_init(&mbar_ag, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
#pragma unroll
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
  // row fo of WF_l: 1.0 in the bias column -> the next layer's constant one
  if (tid < H) {
#pragma unroll
    for (int l = 0; l < H; ++l)
      if (tid == l)
        fs[A.wf_off[l] + wk_index(A.fo[l], A.fi[l], A.KP[l])] = 1.0f;
  }
  __syncthreads();
  for (int f4 = lo4 + tid; f4 < hi4; f4 += FT_THREADS)
    own_w[f4 - lo4] = reinterpret_cast<const float4*>(WFall)[f4];
  if (warp == 0) {
    asm volatile(
        "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
        ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"
                 ::: "memory");
  }
  tc_fence_before();
  fence_async_smem();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // small per-step scratch, in the dynamic shared memory for the same reason
  // (float offsets from scr_a): partial dot products of the output layer
  // [4][FT_ROWS], partial losses [4], Adam step size, minibatch loss (sum /
  // this CTA's), source rows of the next minibatch (32 x s64)
  const uint32_t scr_a = smem_base + 4u * (uint32_t)A.scr_off;
  constexpr uint32_t SC_PART = 0, SC_SQ = 4 * 4 * FT_ROWS,
                     SC_LR = SC_SQ + 16, SC_LOSS = SC_LR + 4,
                     SC_BSQ = SC_LOSS + 4, SC_SRC = SC_BSQ + 16;   // bytes
  uint32_t mbar_a, mbar_ag_a;              // pinned like smem_base
  asm volatile("{\n\t.reg .u64 t;\n\tcvta.to.shared.u64 t, %1;\n\t"
               "cvt.u32.u64 %0, t;\n\t}" : "=r"(mbar_a) : "l"(&mbar));
  asm volatile("{\n\t.reg .u64 t;\n\tcvta.to.shared.u64 t, %1;\n\t"
               "cvt.u32.u64 %0, t;\n\t}" : "=r"(mbar_ag_a) : "l"(&mbar_ag));
  uint32_t phase = 0, phase_ag = 0;

  // Rows of the minibatch are SPREAD over the four TMEM lane quadrants
  // (slot s = 4 * lane + quadrant, lanes 0..7 of every warp): a warp can only
  // reach the 32 TMEM lanes of its quadrant (warp % 4), and this way all
  // sixteen warps -- all four SM sub-partitions -- share the epilogues.  The
  // slot is the K index of the transposed copies.
  const int qd = warp & 3;               // TMEM lanes 32 qd .. 32 qd + 31
  const int part = warp >> 2;            // 0..3: which quarter of the columns
  const bool has_slot = lane < 8;
  const int slot = 4 * lane + qd;        // (meaningful when has_slot)
  const uint32_t tq = tmem + ((uint32_t)(qd * 32) << 16);
  // shared address of (row 0, this thread's slot) of a transposed operand at
  // byte offset 0: row k = c + q of a chunk is (c / 8) * SBO + 16 q further
  const uint32_t tr_base = smem_base + (uint32_t)tr_byte(0, slot);
  const uint32_t wout_a = smem_base + 4u * (uint32_t)A.wout_off;

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
  auto issue_ss = [&](int d_col, int a_byte, int b_byte, int n) {
    const uint64_t da = smem_desc(smem_base + (uint32_t)a_byte, FT_LBO, FT_SBO);
    const uint64_t db = smem_desc(smem_base + (uint32_t)b_byte, FT_LBO, FT_SBO);
    const uint32_t id = idesc_tf32(n);
    for (int s = 0; s < FT_ROWS / 8; ++s)      // K = 32 slots: 4 steps of 8
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
    mbar_wait_a(mbar_a, phase);
    phase ^= 1u;
    tc_fence_after();
  };
  // sum over the eight row lanes of a warp (the other lanes hold zeros)
  auto sum8 = [&](float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
  };

  const long long M = A.m;
  const int n_batches = (int)((M + A.batch - 1) / A.batch);
  const int xw = A.KP[0] + 1;            // floats per prefetched row (x | y)
  float best_loss = INFINITY, last_loss = 0.f;
  int no_improve = 0, epoch = 0;
  long long t_adam = 0;
  int step_no = 0;
#ifdef NB200_FIT_PROF
  // stage clocks of one Adam step (CTA 0, thread 0), printed from the device:
  // build with NB200_EXTRA_FLAGS=-DNB200_FIT_PROF, never in the product
  long long stamp[30];
  int n_stamp = 0;
#define FT_STAMP() do { if (tid == 0 && blockIdx.x == 0 && step_no == 300 && \
                            n_stamp < 30) stamp[n_stamp++] = clock64(); } while (0)
#else
#define FT_STAMP() do {} while (0)
#endif

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
    // rows [first, first + count) of minibatch bi_ that this CTA works on
    auto my_rows = [&](int bi_, long long& first, int& count) {
      const long long lo_ = (long long)bi_ * A.batch;
      const int bn_ = (int)min((long long)A.batch, M - lo_);
      const int per_ = (bn_ + FT_CLUSTER - 1) / FT_CLUSTER;
      const int a_ = min(bn_, crank * per_), b_ = min(bn_, a_ + per_);
      first = lo_ + a_;
      count = b_ - a_;
    };
    // x | y of this CTA's rows of minibatch bi_ -> buffer (bi_ & 1), as 4-byte
    // cp.async copies that complete in the background
    auto prefetch = [&](int bi_, bool src_ready) {
      long long first; int count;
      my_rows(bi_, first, count);
      const int per_row = A.d + 1;
      for (int e = tid; e < count * per_row; e += FT_THREADS) {
        const int s_ = e / per_row, k = e - s_ * per_row;
        const long long src_ = src_ready ? lds_s64(scr_a + SC_SRC + 8u * (uint32_t)s_)
                                         : (pa * (first + s_) + pb) % M;
        const float* g = k < A.d ? x + src_ * A.d + k : y + src_;
        const uint32_t dst =
            smem_base + 4u * (uint32_t)(A.misc_off + (bi_ & 1) * FT_ROWS * xw +
                                        s_ * xw + (k < A.d ? k : A.KP[0]));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
                     ::"r"(dst), "l"(g) : "memory");
      }
    };
    // the first minibatch of the epoch is fetched synchronously
    prefetch(0, false);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();

    float epoch_loss = 0.f;
    for (int bi = 0; bi < n_batches; ++bi) {
      const long long b_lo = (long long)bi * A.batch;
      const int bn = (int)min((long long)A.batch, M - b_lo);
      long long first_row; int R;                // <= FT_ROWS rows
      my_rows(bi, first_row, R);
      const bool row_ok = has_slot && slot < R;
      const bool more = bi + 1 < n_batches;
      FT_STAMP();                                     // 0: step start
      // ---- a_0 = (x | 1): TMEM (A operand) and its transpose -----------------
      // (unconditional loads of this thread's prefetched row, then selects)
      const uint32_t xrow =
          smem_base + 4u * (uint32_t)(A.misc_off + (bi & 1) * FT_ROWS * xw +
                                      (has_slot ? slot : 0) * xw);
      const float y_ld = lds_f32(xrow + 4u * (uint32_t)A.KP[0]);
      const float y_row = row_ok ? y_ld : 0.f;
      {
        const int kp = A.KP[0];
        for (int c = part * 8; c < kp; c += 32) {
          uint32_t v[8];
          const uint32_t ta =
              tr_base + (uint32_t)(A.at_off[0] + (c >> 3) * FT_SBO);
          float f[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) f[q] = lds_f32(xrow + 4u * (uint32_t)(c + q));
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int k = c + q;
            f[q] = !row_ok ? 0.f
                           : (k < A.d ? tf32_rn(f[q]) : (k == A.d ? 1.0f : 0.f));
            v[q] = __float_as_uint(f[q]);
          }
          if (has_slot) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (c + q <= A.d) sts_f32(ta + 16 * q, f[q]);
          }
          tmem_st8(tq + (uint32_t)(A.a_col[0] + c), v);
        }
      }
      FT_STAMP();                                     // (0a: a_0 written)
      publish();
      FT_STAMP();                                     // 1: a_0 published
      // in the shadow of the first MMA: source rows of the NEXT minibatch (one
      // 64-bit modulo each, 32 lanes) and this step's Adam step size
      if (more && tid < FT_ROWS) {
        long long nfirst; int ncount;
        my_rows(bi + 1, nfirst, ncount);
        if (tid < ncount)
          sts_s64(scr_a + SC_SRC + 8u * (uint32_t)tid,
                  (pa * (nfirst + tid) + pb) % M);
      }
      if (tid == 64) {
        const float b1t = powf(A.beta1, (float)(t_adam + 1));
        const float b2t = powf(A.beta2, (float)(t_adam + 1));
        sts_f32(scr_a + SC_LR, A.lr * sqrtf(1.f - b2t) / (1.f - b1t));
      }
      // ---- forward ---------------------------------------------------------------
      // (rows beyond the minibatch carry finite junk from here on: their
      // deltas are zero, so they add nothing to any gradient)
#pragma unroll
      for (int l = 0; l < H; ++l) {
        if (tid == 0) {
          tc_fence_after();
          issue_ts(A.a_col[l + 1], A.a_col[l], A.NP[l], A.KP[l] >> 3,
                   A.wf_off[l], A.KP[l]);
          mma_commit_a(mbar_a);
        }
        await();
        if (l == 0) FT_STAMP();                       // (2a: MMA of L0 done)
        {
          // ReLU -> a_l+1 in place, and its transpose.  The constant-one
          // column (index fo) comes out of the MMA itself: WF_l has a row fo
          // whose only non-zero weight is a 1.0 in the bias column (its
          // gradient is identically zero, Adam never moves it), and the rows
          // beyond it are zero -- so every chunk is treated alike.  A warp's
          // chunks (8 columns every 32) are loaded two at a time before the
          // one wait.
          const int np = A.NP[l];
          // rows of AT_l+1 (the last one feeds the output layer's gradient)
          const int n_tr = l + 1 < H ? A.KP[l + 1] : (fo_last + 8) & ~7;
          for (int c0 = part * 8; c0 < np; c0 += 64) {
            uint32_t v[2][8];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int c = c0 + 32 * j;
              if (c < np) tmem_ld8(tq + (uint32_t)(A.a_col[l + 1] + c), v[j]);
            }
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int c = c0 + 32 * j;
              if (c >= np) continue;
#pragma unroll
              for (int q = 0; q < 8; ++q)
                v[j][q] = tf32_bits(fmaxf(__uint_as_float(v[j][q]), 0.f));
              if (has_slot && c < n_tr) {
                const uint32_t ta =
                    tr_base + (uint32_t)(A.at_off[l + 1] + (c >> 3) * FT_SBO);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  sts_f32(ta + 16 * q, __uint_as_float(v[j][q]));
              }
              tmem_st8(tq + (uint32_t)(A.a_col[l + 1] + c), v[j]);
            }
          }
        }
        if (l == 0) FT_STAMP();                       // (2b: epilogue of L0)
#ifdef NB200_FIT_PROF
        if (l == 0) {                                 // publish(), taken apart
          tmem_wait_st();
          FT_STAMP();
          tc_fence_before();
          fence_async_smem();
          FT_STAMP();
          __syncthreads();
        } else
#endif
        publish();
        // (s_src is visible now: fetch the next minibatch in the background)
        if (l == 0 && more) prefetch(bi + 1, true);
        FT_STAMP();                                   // 2..4: forward layers
      }
      // ---- output layer, loss, e_H-1 (CUDA cores; fan_out 1) ---------------------
      // warp (qd, part) handles columns [8 part, 8 part + 8) + 32 j of a_H for
      // the rows of its quadrant; partial dot products meet in shared memory
      // warp (qd, part) handles columns [8 part, 8 part + 8) + 32 j of a_H for
      // the rows of its quadrant; the partial dot products meet in shared
      // memory.  (One warp per quadrant doing all columns, without the
      // exchange, measured 2.5x slower.)
      // OC: 8-column chunks of a warp (1 for a last hidden layer of <= 31
      // units, the reference's default; else up to 4)
      auto output_stage = [&](auto oc_tag) {
        constexpr int OC = decltype(oc_tag)::value;
        const int np = A.NP[H - 1], kb = A.KB[H - 1];
        uint32_t v[OC][8];
#pragma unroll
        for (int j = 0; j < OC; ++j) {
          const int c = part * 8 + 32 * j;
          if (c < np) {
            tmem_ld8(tq + (uint32_t)(A.a_col[H] + c), v[j]);
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) v[j][q] = 0u;
          }
        }
        tmem_wait_ld();
        FT_STAMP();                                   // (5a: a_H in registers)
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < OC; ++j) {
          const int c = part * 8 + 32 * j;
          if (c >= np) continue;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (c + q < fo_last)
              acc = fmaf(__uint_as_float(v[j][q]),
                         lds_f32(wout_a + 4u * (uint32_t)(c + q)), acc);
        }
        if (has_slot)
          sts_f32(scr_a + SC_PART + 4u * (uint32_t)(part * FT_ROWS + slot), acc);
        __syncthreads();
        FT_STAMP();                                   // (5b: partial sums met)
        const uint32_t pa_ = scr_a + SC_PART + 4u * (uint32_t)(has_slot ? slot : 0);
        const float yhat = lds_f32(pa_) + lds_f32(pa_ + 4 * FT_ROWS) +
                           lds_f32(pa_ + 8 * FT_ROWS) +
                           lds_f32(pa_ + 12 * FT_ROWS) +
                           lds_f32(wout_a + 4u * (uint32_t)fo_last);
        const float diff = row_ok ? yhat - y_row : 0.f;
        const float e_out = diff / (float)bn;  // (y - t) / batch of this row
        if (part == 0) {
          const float sq = sum8(diff * diff);
          if (lane == 0) sts_f32(scr_a + SC_SQ + 4u * (uint32_t)qd, sq);
          // e_H (one unit) transposed: the gradient of the output layer is one
          // more tensor-core product, (e_H)^T (a_H | 1)
          if (has_slot)
            sts_f32(tr_base + (uint32_t)A.et_off[H], tf32_rn(e_out));
        }
        FT_STAMP();                                   // (5c: loss, e_out)
        // e_H-1 = e_out W_H (a_H > 0)
#pragma unroll
        for (int j = 0; j < OC; ++j) {
          const int c = part * 8 + 32 * j;
          if (c >= max(kb, np)) continue;
          const uint32_t ta =
              tr_base + (uint32_t)(A.et_off[H - 1] + (c >> 3) * FT_SBO);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int k = c + q;
            const float a = k < fo_last ? __uint_as_float(v[j][q]) : 0.f;
            v[j][q] = a > 0.f
                          ? to_tf32(e_out * lds_f32(wout_a + 4u * (uint32_t)(
                                                    k < fo_last ? k : 0)))
                          : 0u;
          }
          if (has_slot && c < np) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              sts_f32(ta + 16 * q, __uint_as_float(v[j][q]));
          }
          if (c < kb) tmem_st8(tq + (uint32_t)(A.e_col[H - 1] + c), v[j]);
        }
        FT_STAMP();                                   // (5d: e_H-1 written)
      };
      if (A.NP[H - 1] <= 32) output_stage(std::integral_constant<int, 1>{});
      else output_stage(std::integral_constant<int, FT_MAX_W / 32>{});
      publish();
      FT_STAMP();                                     // 5: output layer, e_H-1
      // ---- deltas of the earlier hidden layers -----------------------------------
#pragma unroll
      for (int l = H - 2; l >= 0; --l) {
        if (tid == 0) {
          tc_fence_after();
          // e_l = e_l+1 . W_l+1^T : N = round16(fo_l) = NB[l+1], K = KB[l+1]
          issue_ts(A.e_col[l], A.e_col[l + 1], A.NB[l + 1], A.KB[l + 1] >> 3,
                   A.wb_off[l + 1], A.KB[l + 1]);
          mma_commit_a(mbar_a);
        }
        await();
        {
          const int nb = A.NB[l + 1];
          for (int c0 = part * 8; c0 < nb; c0 += 64) {
            uint32_t v[2][8], a[2][8];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int c = c0 + 32 * j;
              if (c < nb) {
                tmem_ld8(tq + (uint32_t)(A.e_col[l] + c), v[j]);
                tmem_ld8(tq + (uint32_t)(A.a_col[l + 1] + c), a[j]);
              }
            }
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int c = c0 + 32 * j;
              if (c >= nb) continue;
              const uint32_t ta =
                  tr_base + (uint32_t)(A.et_off[l] + (c >> 3) * FT_SBO);
              // (rows beyond the minibatch: e_l+1 = 0, so e_l = 0 already;
              // columns beyond fo: zero rows of WB, so the product is 0)
#pragma unroll
              for (int q = 0; q < 8; ++q)
                v[j][q] = __uint_as_float(a[j][q]) > 0.f
                              ? to_tf32(__uint_as_float(v[j][q])) : 0u;
              if (has_slot) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  sts_f32(ta + 16 * q, __uint_as_float(v[j][q]));
              }
              if (l > 0) tmem_st8(tq + (uint32_t)(A.e_col[l] + c), v[j]);
            }
          }
        }
        publish();
        FT_STAMP();                                   // 6..7: deltas
      }
      // ---- gradients gW_l^T = e_l^T (a_l | 1): one MMA group per layer -----------
      float batch_sq = 0.f;
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int l = 0; l <= H; ++l)
          issue_ss(A.g_col[l], A.et_off[l], A.at_off[l], A.NG[l]);
        mma_commit_a(mbar_a);
        batch_sq = lds_f32(scr_a + SC_SQ) + lds_f32(scr_a + SC_SQ + 4) +
                   lds_f32(scr_a + SC_SQ + 8) + lds_f32(scr_a + SC_SQ + 12);
        sts_f32(scr_a + SC_BSQ, batch_sq);
      }
      await();
      FT_STAMP();                                     // 8: gradient MMAs done
      {
        // D[lane = output o][col = input i (fi = the bias)] -> the gradient,
        // in place of WF_l: columns c .. c + 7 are two 16-byte chunks
        const int o = qd * 32 + lane;
#pragma unroll
        for (int l = 0; l < H; ++l) {
          const int kp = A.KP[l], np = A.NP[l];
          if (qd * 32 >= np) continue;      // this quadrant holds no unit
          for (int c0 = part * 8; c0 < kp; c0 += 128) {
            uint32_t v[4][8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = c0 + 32 * j;
              if (c < kp) tmem_ld8(tq + (uint32_t)(A.g_col[l] + c), v[j]);
            }
            tmem_wait_ld();
            if (o < np) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int c = c0 + 32 * j;
                if (c >= kp) continue;
                const uint32_t dst =
                    smem_base +
                    4u * (uint32_t)(A.wf_off[l] + (o >> 3) * (kp * 8) +
                                    (c >> 2) * 32 + (o & 7) * 4);
                sts_v4(dst, make_float4(__uint_as_float(v[j][0]),
                                        __uint_as_float(v[j][1]),
                                        __uint_as_float(v[j][2]),
                                        __uint_as_float(v[j][3])));
                sts_v4(dst + 128, make_float4(__uint_as_float(v[j][4]),
                                              __uint_as_float(v[j][5]),
                                              __uint_as_float(v[j][6]),
                                              __uint_as_float(v[j][7])));
              }
            }
          }
        }
      }
      if (qd == 0) {
        // lane 0 of the output layer's tile: gW_H (fo_last values), then gb
        for (int c = part * 8; c <= fo_last; c += 32) {
          uint32_t v[8];
          tmem_ld8(tq + (uint32_t)(A.g_col[H] + c), v);
          tmem_wait_ld();
          if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (c + q <= fo_last) gout[c + q] = __uint_as_float(v[q]);
          }
        }
      }
      FT_STAMP();                                     // 9: gradient in WF
      // ---- reduce-scatter of the gradient over the cluster (128-bit distributed-
      // shared-memory loads), Adam on this CTA's share of the WF index space -----
      asm volatile("cp.async.wait_all;" ::: "memory");
      tc_fence_before();
      cluster.sync();
      FT_STAMP();                                     // 10: cluster barrier 1
      t_adam += 1;
      {
        const float lr_t = lds_f32(scr_a + SC_LR);
        const float c1 = 1.f - A.beta1, c2 = 1.f - A.beta2;
        const uint32_t wf_a = smem_base + 4u * (uint32_t)A.wf_off[0];
        const uint32_t own_a = smem_base + 4u * (uint32_t)A.own_off;
        for (int f4 = lo4 + tid; f4 < hi4; f4 += FT_THREADS) {
          float4 gq[FT_CLUSTER];
#pragma unroll
          for (int q = 0; q < FT_CLUSTER; ++q)
            gq[q] = ldc_v4(mapa_u32(wf_a + 16u * (uint32_t)f4, q));
          const uint32_t mine = own_a + 16u * (uint32_t)(f4 - lo4);
          float4 mq = lds_v4(mine);
          float4 vq = lds_v4(mine + 16u * (uint32_t)A.share4);
          float4 w = lds_v4(mine + 32u * (uint32_t)A.share4);
          float4 g = gq[0];
#pragma unroll
          for (int q = 1; q < FT_CLUSTER; ++q) g = f4_add(g, gq[q]);
          mq.x = A.beta1 * mq.x + c1 * g.x; vq.x = A.beta2 * vq.x + c2 * g.x * g.x;
          mq.y = A.beta1 * mq.y + c1 * g.y; vq.y = A.beta2 * vq.y + c2 * g.y * g.y;
          mq.z = A.beta1 * mq.z + c1 * g.z; vq.z = A.beta2 * vq.z + c2 * g.z * g.z;
          mq.w = A.beta1 * mq.w + c1 * g.w; vq.w = A.beta2 * vq.w + c2 * g.w * g.w;
          w.x -= lr_t * mq.x / (sqrtf(vq.x) + A.eps);
          w.y -= lr_t * mq.y / (sqrtf(vq.y) + A.eps);
          w.z -= lr_t * mq.z / (sqrtf(vq.z) + A.eps);
          w.w -= lr_t * mq.w / (sqrtf(vq.w) + A.eps);
          sts_v4(mine, mq);
          sts_v4(mine + 16u * (uint32_t)A.share4, vq);
          sts_v4(mine + 32u * (uint32_t)A.share4, w);
        }
        // the output layer (fo_last + 1 parameters), replicated on every CTA
        if (FT_THREADS - 1 - tid <= fo_last) {
          const int k = FT_THREADS - 1 - tid;
          float g = 0.f;
#pragma unroll
          for (int q = 0; q < FT_CLUSTER; ++q)
            g += cluster.map_shared_rank(gout, q)[k];
          const float mq = A.beta1 * mout[k] + c1 * g;
          const float vq = A.beta2 * mout[FT_MAX_W + 1 + k] + c2 * g * g;
          mout[k] = mq;
          mout[FT_MAX_W + 1 + k] = vq;
          wout[k] -= lr_t * mq / (sqrtf(vq) + A.eps);
        }
        // the minibatch loss: one warp fetches the eight partial sums
        if (warp == 8) {
          float t = 0.f;
          if (lane < FT_CLUSTER)
            asm volatile("ld.shared::cluster.f32 %0, [%1];"
                         : "=f"(t) : "r"(mapa_u32(scr_a + SC_BSQ, lane))
                         : "memory");
          t = sum8(t);
          if (lane == 0) sts_f32(scr_a + SC_LOSS, t);
        }
      }
      FT_STAMP();                                     // 11: reduce + Adam
      // ---- all-gather: this CTA's share of the new weights goes into the WF of
      // every CTA as one bulk copy each (TMA, shared::cta -> shared::cluster),
      // completing on the destination's mbarrier.  Nobody else reads or writes
      // this share of anybody's WF, and a CTA that has received all eight
      // shares knows that every peer is past its reduce: no cluster barrier.
      fence_async_smem();                  // own_w: generic -> async proxy
      __syncthreads();
      if (tid == 0) {
        mbar_expect_tx_a(mbar_ag_a, (uint32_t)A.wf_floats * 4u);
        const uint32_t src =
            smem_base + 4u * (uint32_t)A.own_off + 32u * (uint32_t)A.share4;
        const uint32_t dst =
            smem_base + 4u * (uint32_t)A.wf_off[0] + 16u * (uint32_t)lo4;
        const uint32_t bytes = 16u * (uint32_t)(hi4 - lo4);
        const uint32_t bar = mbar_ag_a;
#pragma unroll
        for (int q = 0; q < FT_CLUSTER; ++q) {
          uint32_t rdst, rbar;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                       : "=r"(rdst) : "r"(dst), "r"(q));
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                       : "=r"(rbar) : "r"(bar), "r"(q));
          asm volatile(
              "cp.async.bulk.shared::cluster.shared::cta.mbarrier::"
              "complete_tx::bytes [%0], [%1], %2, [%3];"
              ::"r"(rdst), "r"(src), "r"(bytes), "r"(rbar) : "memory");
        }
      }
      mbar_wait_a(mbar_ag_a, phase_ag);
      phase_ag ^= 1u;
      FT_STAMP();                                     // 12: all-gather landed
      refresh_wb();
      // (WB was written through the generic proxy; the MMAs read it through
      // the async proxy)
      fence_async_smem();
      tc_fence_after();
      __syncthreads();
      batch_sq = lds_f32(scr_a + SC_LOSS);
      FT_STAMP();                                     // 13: WB refreshed
#ifdef NB200_FIT_PROF
      if (tid == 0 && blockIdx.x == 0 && step_no == 300) {
        printf("fit_tc stage clocks:");
        for (int q = 1; q < n_stamp; ++q)
          printf(" %d:%lld", q, stamp[q] - stamp[q - 1]);
        printf("  total %lld\n", stamp[n_stamp - 1] - stamp[0]);
      }
#endif
      epoch_loss += 0.5f * batch_sq;
      step_no += 1;
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
  int fl = 0, col = 0;                       // shared-memory floats, TMEM cols
  for (int l = 0; l < H; ++l) {
    A.fi[l] = sizes[l]; A.fo[l] = sizes[l + 1];
    if (A.fi[l] + 1 > 128 || A.fo[l] + 1 > FT_MAX_W) return false;
    A.KP[l] = r8(A.fi[l] + 1); A.NP[l] = r16(A.fo[l] + 1);
    A.KB[l] = r8(A.fo[l]); A.NB[l] = r16(A.fi[l]);
    A.NG[l] = r16(A.fi[l] + 1);
    const unsigned long long dv = 8ull * A.KP[l];
    A.magic[l] = (unsigned int)(((1ull << 32) + dv - 1) / dv);
  }
  // transposed operands first: the gradient MMAs read 128 rows (16 groups) of
  // every ET_l and NG_l rows of AT_l; what lies behind the rows that exist is
  // other, readable data
  int bytes = 0;
  for (int l = 0; l < H; ++l) {
    A.at_off[l] = bytes; bytes += (A.fi[l] + 1 + 7) / 8 * FT_SBO;
  }
  A.at_off[H] = bytes; bytes += (sizes[H] + 1 + 7) / 8 * FT_SBO;
  for (int l = 0; l < H; ++l) {
    A.et_off[l] = bytes; bytes += A.NP[l] / 8 * FT_SBO;
  }
  A.et_off[H] = bytes; bytes += FT_SBO;           // one unit: one 8-row group
  A.NG[H] = r16(sizes[H] + 1);
  bytes = (bytes + 127) / 128 * 128;
  fl = bytes / 4;
  // WF_0 | WF_1 | ... contiguous: the index space of the reduce-scatter
  for (int l = 0; l < H; ++l) {
    A.wf_off[l] = fl; fl += A.NP[l] * A.KP[l];     // (a multiple of 128)
  }
  A.wf_floats = fl - A.wf_off[0];
  A.share4 = (A.wf_floats / 4 + FT_CLUSTER - 1) / FT_CLUSTER;
  for (int l = 1; l < H; ++l) {
    A.wb_off[l] = fl; fl += A.NB[l] * A.KB[l];
    fl = (fl + 31) / 32 * 32;
  }
  A.own_off = fl; fl += 3 * 4 * A.share4;
  A.wout_off = fl; fl += FT_MAX_W + 1;
  A.gout_off = fl; fl += FT_MAX_W + 1;
  A.mout_off = fl; fl += 2 * (FT_MAX_W + 1);
  fl = (fl + 31) / 32 * 32;
  // next minibatch's rows, prefetched in the background: two buffers of
  // FT_ROWS x (KP_0 inputs + the target)
  A.misc_off = fl; fl += 2 * FT_ROWS * (A.KP[0] + 1);
  fl = (fl + 31) / 32 * 32;
  A.scr_off = fl; fl += 4 * FT_ROWS + 12 + 2 * FT_ROWS;
  fl = (fl + 31) / 32 * 32;
  A.smem_bytes = fl * 4;
  // (the over-reads of the gradient operands must stay inside: small
  // networks get the slack as padding)
  if (A.et_off[H] + 16 * FT_SBO > A.smem_bytes)
    A.smem_bytes = (A.et_off[H] + 16 * FT_SBO + 127) / 128 * 128;
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
  // the gradient tiles lie over the activations and deltas, dead by then
  col = 0;
  for (int l = 0; l <= H; ++l) { A.g_col[l] = col; col += r32(A.NG[l]); }
  if (col > 512) return false;
  *out = A;
  return true;
}

int launch_fit_tc(FitTcArgs A, const float* x32, const float* y32,
                  float* moments, int n_net, double* weights_out,
                  int* n_iter_out, double* loss_out, cudaStream_t st) {
#define NB200_FIT_TC_LAUNCH(HH)                                               \
  do {                                                                        \
    NB_CUDA(cudaFuncSetAttribute(k_mlp_fit_tc<HH>,                            \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                 A.smem_bytes));                              \
    k_mlp_fit_tc<HH><<<n_net * FT_CLUSTER, FT_THREADS, A.smem_bytes, st>>>(   \
        A, x32, y32, moments, weights_out, n_iter_out, loss_out);             \
  } while (0)
  if (A.H == 1) NB200_FIT_TC_LAUNCH(1);
  else if (A.H == 2) NB200_FIT_TC_LAUNCH(2);
  else NB200_FIT_TC_LAUNCH(3);
#undef NB200_FIT_TC_LAUNCH
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
