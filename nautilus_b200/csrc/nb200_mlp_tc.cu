// Emulator forward pass on the 5th-generation tensor cores (tcgen05, TMEM).
//
// Replaces NeuralNetworkEmulator.predict (nautilus/neural.py:114-116), i.e.
// n_networks x sklearn MLPRegressor._forward_pass_fast
// (sklearn/neural_network/_multilayer_perceptron.py:189-224), for the one
// genuinely dense contraction on the hot path.
//
// Mapping (DESIGN.md "emulator kernel"):
//   * one persistent CTA per SM, 2 tile groups x 4 warps; a group owns a tile
//     of 128 points, thread r of the group == point r == TMEM lane r;
//   * all weights of all networks (tf32, K-major core-matrix layout, zero
//     padded) are fetched ONCE per CTA by a TMA bulk copy (cp.async.bulk ->
//     UBLKCP) and stay resident in shared memory (~193 KB for 4 x 30-100-50-20-1);
//   * activations never touch shared or global memory: the standardised
//     input row is written to TMEM with tcgen05.st, every hidden layer is
//     D[128 x N] = A[128 x K] (TMEM) . W^T (smem) with tcgen05.mma kind::tf32
//     (fp32 accumulate), the epilogue reads D with tcgen05.ld, adds the bias,
//     applies ReLU, rounds to tf32 and writes it back IN PLACE as the next
//     layer's A operand; the last layer (fan_out 1) is a dot product in the
//     epilogue registers;
//   * the two groups of a CTA interleave, so one group's MMAs run under the
//     other group's epilogue.
// Arithmetic is tf32 x tf32 -> fp32, so scores differ from the fp64 path by
// ~1e-3; membership near the threshold can flip (rate reported by the tests),
// which is why the fp64 kernel remains the parity mode.
#include "nb200_device.cuh"
#include "nb200_tc.cuh"

namespace nb200 {

constexpr int TC_MAX_HID = 4;
constexpr int TC_HDR_WORDS = 32;
constexpr int TC_GROUPS = 2;            // tiles in flight per CTA
constexpr int TC_GROUP_THREADS = 256;   // two threads per row of a tile
constexpr int TC_COLS_PER_GROUP = 256;

struct TcHeader {            // int32[32] in the meta tail (_pack.py:pack_tc)
  int magic, n_net, n_hid, d;   // magic = 0x7F32 | (tile groups per CTA << 16)
  int k0p, net_stride, total_floats, a0_col;
  int np[TC_MAX_HID];        // padded fan_out of hidden layer l (mult. of 16)
  int kp[TC_MAX_HID];        // padded fan_in  of hidden layer l (mult. of 8)
  int w_off[TC_MAX_HID];     // float offsets inside a network block
  int b_off[TC_MAX_HID];
  int d_col[TC_MAX_HID];     // TMEM column of layer l's accumulator
  int w_out_off, b_out_off;
  int thr_lo, thr_hi;        // bits of the fp64 threshold score_predict_min-1e-9
};
static_assert(sizeof(TcHeader) == TC_HDR_WORDS * 4, "header size");

// Optional fused tail: the per-block partials of the shell sums (nb200_stats)
// over the likelihoods k_front already evaluated (log_l, NaN for the rows it
// rejected), so that a cycle without later bounds is front kernel -> this
// kernel -> one tiny final reduction.  Rows this kernel rejects get NaN.
struct TcTail {
  double* log_l;
  StatPartial* partial;     // nullptr: tail disabled
  double log_l_min;
};

// Debug timeline (clock64 stamps of the leader of group 0 in CTA 0), read
// back with nb200_debug_read(); compiled in only with -DNB200_TIMELINE.
#ifdef NB200_TIMELINE
__device__ long long g_tl[512];
__device__ int g_tl_n;
#define TL_STAMP(tag)                                                   \
  do {                                                                  \
    if (blockIdx.x == 0 && threadIdx.x == 0 && g_tl_n < 510) {          \
      g_tl[g_tl_n++] = (long long)(tag);                                \
      g_tl[g_tl_n++] = clock64();                                       \
    }                                                                   \
  } while (0)
#else
#define TL_STAMP(tag) do {} while (0)
#endif

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_GROUPS * TC_GROUP_THREADS, 1)
k_mlp_tf32(const TcHeader h, const float* __restrict__ blob,
           const float* __restrict__ xs32, const uint8_t* __restrict__ mask,
           int64_t n, double* __restrict__ score_out,
           uint8_t* __restrict__ passf, uint8_t* __restrict__ code,
           const TcTail tail) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t wbar;
  __shared__ uint64_t mbar[TC_GROUPS * 3];
  __shared__ uint32_t tmem_slot;
  __shared__ float part[TC_GROUPS][128];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int g = tid >> 8;            // tile group (8 warps)
  const int r = tid & 127;           // row in tile == TMEM lane
  const int hf = (tid >> 7) & 1;     // which half of the columns this thread
                                     // handles (warps w and w+4 share lanes)

  const double thr = __hiloint2double(h.thr_hi, h.thr_lo);
  // architectures that need more than 256 TMEM columns run ONE tile group per
  // CTA (group 1 idles); otherwise two groups interleave
  const int n_groups = h.magic >> 16;
  if (tid == 0) {
    mbar_init(&wbar, 1);
    for (int q = 0; q < TC_GROUPS * 3; ++q) mbar_init(&mbar[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float* wsm = (float*)smem;
  if (warp == 0) {
    asm volatile(
        "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
        ::"r"(smem_u32(&tmem_slot)), "r"(TC_GROUPS * TC_COLS_PER_GROUP)
        : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"
                 ::: "memory");
  }
  if (tid == 32) {   // one thread of warp 1 feeds the weights by TMA bulk copy
    const uint32_t bytes = (uint32_t)h.total_floats * 4u;
    mbar_expect_tx(&wbar, bytes);
    const char* src = (const char*)blob;
    uint32_t done = 0;
    while (done < bytes) {       // chunks: keep each copy <= 64 KB
      const uint32_t c = min(bytes - done, 65536u);
      bulk_g2s(smem + done, src + done, c, &wbar);
      done += c;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot + (uint32_t)(g * TC_COLS_PER_GROUP);
  // Warp-uniform copies for the MMA issuer: values that come out of a
  // shuffle are uniform to the compiler, so the operands of tcgen05.mma live
  // in uniform registers instead of being moved there one by one (R2UR) --
  // the timeline showed ~200 cycles per MMA issue before this.
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const bool issuer_warp = (warp_u & 7) == 0;          // first warp of a group
  const uint32_t tmem_base_u =
      __shfl_sync(0xffffffffu, tmem_slot, 0) +
      (uint32_t)((warp_u >> 3) * TC_COLS_PER_GROUP);
  const uint32_t wsm_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
  mbar_wait(&wbar, 0);

  // ---- MMA issue table: everything tcgen05.mma needs per layer, computed
  // once (the timeline showed 600-900 cycles per issue when the descriptors
  // were rebuilt from the header on the critical path) ----------------------
  __shared__ uint4 mma_tab[TC_MAX_HID];       // {d_col, a_col, idesc, k steps}
  __shared__ uint64_t desc_tab[TC_MAX_HID];   // weight descriptor of network 0
  if (tid < h.n_hid) {
    const int l = tid;
    mma_tab[l] = make_uint4(
        (uint32_t)h.d_col[l],
        (uint32_t)(l == 0 ? h.a0_col : h.d_col[l - 1]),
        idesc_tf32(h.np[l]), (uint32_t)(h.kp[l] >> 3));
    desc_tab[l] = smem_desc(smem_u32(smem) + 4u * (uint32_t)h.w_off[l], 128u,
                            (uint32_t)h.kp[l] * 32u);
  }
  __syncthreads();
  // descriptor address field counts 16-byte units: + net * net_stride floats
  const uint32_t net_step16 = (uint32_t)h.net_stride >> 2;

  Lse lse_acc;
  lse_acc.init();
  int c_rej0 = 0, c_rej1 = 0, c_rej2 = 0, c_rej3 = 0, c_in = 0, c_upd = 0,
      c_raw = 0;

  uint32_t phA = 0, phB = 0, phC = 0;
  const int64_t n_tiles = (n + 127) / 128;
  const int64_t tile_step = (int64_t)gridDim.x * n_groups;
  const int64_t tile0 = g < n_groups ? (int64_t)blockIdx.x * n_groups + g
                                     : n_tiles;
  const bool with_tail = tail.partial != nullptr;
  // The inputs of the NEXT tile are fetched into registers while the current
  // tile runs: this half's 16 columns of the standardised row (k0p <= 32:
  // 4 x 16 B), the candidate flag, and for the fused tail the disposition
  // byte and the likelihood the front kernel already evaluated.  None of the
  // loads depends on another one.
  const bool prefetch = h.k0p <= 32;
  uint4 pre[4];
  uint32_t nx_mask = 0, nx_cd = NB200_CODE_IN_SHELL;
  double nx_ll = 0.0;
  auto fetch = [&](int64_t t) {
    const int64_t rw = t * 128 + r;
    nx_mask = 0;
    if (t < n_tiles && rw < n) {
      nx_mask = mask ? (uint32_t)mask[rw] : 1u;
      if (with_tail && !hf) {
        nx_cd = code[rw];
        nx_ll = tail.log_l[rw];
      }
      if (prefetch) {
        const uint4* src =
            (const uint4*)(xs32 + rw * (int64_t)h.k0p) + hf * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (hf * 16 + q * 4 < h.k0p) pre[q] = __ldg(src + q);
      }
    }
  };
  // standardised input rows of tile t -> TMEM (A operand of layer 0)
  auto stage_a0 = [&](int64_t t, bool active) {
    if (prefetch) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = hf * 16 + q * 8;
        if (c < h.k0p) {
          uint32_t v[8];
          const uint4 a = active ? pre[2 * q] : make_uint4(0, 0, 0, 0);
          const uint4 b = active ? pre[2 * q + 1] : make_uint4(0, 0, 0, 0);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
          v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
          tmem_st8(tmem_base + lane_addr + (uint32_t)(h.a0_col + c), v);
        }
      }
    } else {
      const int64_t rw = t * 128 + r;
      const uint4* src = (const uint4*)(xs32 + rw * (int64_t)h.k0p);
      for (int c = hf * 8; c < h.k0p; c += 16) {
        uint32_t v[8];
        if (active) {
          const uint4 a = __ldg(src + (c >> 2));
          const uint4 b = __ldg(src + (c >> 2) + 1);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
          v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = 0u;
        }
        tmem_st8(tmem_base + lane_addr + (uint32_t)(h.a0_col + c), v);
      }
    }
  };

  // ---- helpers ------------------------------------------------------------
  // leader: issue layer l of network `net`, completion arrives on `bar`
  auto issue = [&](int net, int l, uint64_t* bar) {
    if (issuer_warp) {                 // warp-uniform branch
      tc_fence_after();
      if (elect_one()) {
        const uint4 t = mma_tab[l];
        const uint64_t desc0 =
            desc_tab[l] + (uint64_t)((uint32_t)net * net_step16);
        const uint32_t d_tmem = tmem_base_u + t.x;
        const uint32_t a_tmem = tmem_base_u + t.y;
        const int ks = (int)t.w;
        for (int s = 0; s < ks; ++s) {
          // one K-step = 2 core matrices = 256 B -> +16 in the 16-B units
          // of the descriptor's start-address field
          mma_tf32_ts(d_tmem, a_tmem + (uint32_t)(s * 8),
                      desc0 + (uint64_t)(s * 16), t.z, s > 0 ? 1u : 0u);
        }
        mma_commit(bar);
      }
      __syncwarp();
    }
  };
  // all: ReLU + tf32 rounding of layer l's accumulator, written back in
  // place as the next layer's A operand (the bias was added by the MMA
  // through the constant-one column, see _pack.py:pack_tc)
  auto epi_hidden = [&](int l) {
    tc_fence_after();
    const uint32_t d_addr = tmem_base + lane_addr + (uint32_t)h.d_col[l];
    // 32 columns per TMEM load (the accumulator regions are 32-column
    // aligned; pad columns meet zero weights), halves interleaved
    for (int c = hf * 32; c < h.np[l]; c += 64) {
      uint32_t v[32];
      tmem_ld32(d_addr + (uint32_t)c, v);
      tmem_wait_ld();
      // ReLU, then round-half-up to tf32: the MMA reads only the top 19
      // bits, so adding half an ulp of tf32 is the whole rounding
#pragma unroll
      for (int q = 0; q < 32; ++q)
        v[q] = __float_as_uint(fmaxf(__uint_as_float(v[q]), 0.f)) + 0x1000u;
      tmem_st32(d_addr + (uint32_t)c, v);
    }
    tmem_wait_st();
    tc_fence_before();
    group_sync(g);
  };
  // all: last hidden layer, the fan_out-1 output layer folded in registers
  auto epi_last = [&](int net, int l) -> float {
    tc_fence_after();
    const float* wnet = wsm + (size_t)net * h.net_stride;
    const float* wout = wnet + h.w_out_off;
    const uint32_t d_addr = tmem_base + lane_addr + (uint32_t)h.d_col[l];
    float acc = hf ? 0.f : wnet[h.b_out_off];
    for (int c = hf * 16; c < h.np[l]; c += 32) {
      uint32_t v[16];
      tmem_ld16(d_addr + (uint32_t)c, v);
      // the output weights of these 16 columns: four broadcast LDS.128
      const float4* w4 = reinterpret_cast<const float4*>(wout + c);
      const float4 w0 = w4[0], w1 = w4[1], w2 = w4[2], w3 = w4[3];
      const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                           w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
      tmem_wait_ld();
#pragma unroll
      for (int q = 0; q < 16; ++q)
        acc = fmaf(fmaxf(__uint_as_float(v[q]), 0.f), w[q], acc);
    }
    if (hf) part[g][r] = acc;
    // a later MMA overwrites these columns: order the loads before it
    tc_fence_before();
    group_sync(g);
    return hf ? 0.f : acc + part[g][r];
  };
  // half 0: score -> decision -> outputs of one row, and with the fused tail
  // the disposition histogram and the log-sum-exp of the likelihood that
  // k_front evaluated for this row
  auto finalize = [&](int64_t tile, float sum, bool active, uint32_t cd_in,
                      double ll) {
    const int64_t row = tile * 128 + r;
    bool accepted = false;
    if (active) {
      const double score = (double)(sum / (float)h.n_net);
      accepted = score > thr;
      if (score_out) score_out[row] = score;
      if (passf && accepted) passf[row] = 1;
      if (code && !accepted) code[row] = NB200_CODE_NN_REJECT;
    } else if (score_out && row < n) {
      score_out[row] = nan("");
    }
    if (with_tail && row < n) {
      c_raw += 1;
      if (active ? accepted : cd_in == NB200_CODE_IN_SHELL) {
        lse_acc.add(ll);
        c_in += 1;
        if (ll >= tail.log_l_min) c_upd += 1;
      } else {
        if (active) {
          tail.log_l[row] = nan("");      // the front's value is void now
          c_rej2 += 1;
        } else {
          c_rej0 += cd_in == NB200_CODE_CUBE_REJECT;
          c_rej1 += cd_in == NB200_CODE_OVERLAP_REJECT;
          c_rej2 += cd_in == NB200_CODE_NN_REJECT;
          c_rej3 += cd_in == NB200_CODE_EXCLUDED;
        }
      }
    }
  };

  uint64_t* bA = &mbar[g * 3 + 0];
  uint64_t* bB = &mbar[g * 3 + 1];
  uint64_t* bC = &mbar[g * 3 + 2];
  fetch(tile0);
  if (h.n_hid == 3) {
    // One continuous stream of (tile, network) instances per group, two in
    // flight: while the CUDA cores run one instance's epilogue the tensor
    // pipe runs the other's next layer,
    //   L0(i+1) || epilogue1(i),  L2(i) || epilogue0(i+1),
    //   L1(i+1) || epilogue2(i),
    // and the stream does not drain at tile boundaries: the next tile's
    // input rows go to TMEM as soon as the last network of this tile has
    // consumed the old ones.
    int64_t tile = tile0;
    bool have = tile < n_tiles;
    bool cur_active = false, hold_active = false;
    uint32_t cur_cd = NB200_CODE_IN_SHELL, hold_cd = NB200_CODE_IN_SHELL;
    double cur_ll = 0.0, hold_ll = 0.0;
    if (have) {
      TL_STAMP(1);
      cur_active = nx_mask != 0; cur_cd = nx_cd; cur_ll = nx_ll;
      stage_a0(tile, cur_active);
      tmem_wait_st();
      fetch(tile + tile_step);
      tc_fence_before();
      group_sync(g);
      TL_STAMP(2);
      issue(0, 0, bA);
      mbar_wait(bA, phA); phA ^= 1u;
      epi_hidden(0);
      issue(0, 1, bB);
    }
    while (have) {
      const int64_t next_tile = tile + tile_step;
      const bool have_next = next_tile < n_tiles;
      float sum = 0.f;
      for (int net = 0; net < h.n_net; ++net) {
        const bool last = net + 1 == h.n_net;
        const bool more = !last || have_next;
        const int nn = last ? 0 : net + 1;
        TL_STAMP(10);
        mbar_wait(bB, phB); phB ^= 1u;     // D1(i) ready, region 0 free
        TL_STAMP(11);
        if (last && have_next) {
          // every network of this tile has read the input rows: stage the
          // next tile's
          hold_active = nx_mask != 0; hold_cd = nx_cd; hold_ll = nx_ll;
          stage_a0(next_tile, hold_active);
          tmem_wait_st();
          fetch(next_tile + tile_step);
          tc_fence_before();
          group_sync(g);
          TL_STAMP(1);
        }
        if (more) issue(nn, 0, bA);
        TL_STAMP(12);
        epi_hidden(1);
        TL_STAMP(13);
        issue(net, 2, bC);
        TL_STAMP(14);
        if (more) {
          mbar_wait(bA, phA); phA ^= 1u;
          TL_STAMP(15);
          epi_hidden(0);
          TL_STAMP(16);
        }
        mbar_wait(bC, phC); phC ^= 1u;     // D2(i) ready, region 1 free
        TL_STAMP(17);
        if (more) issue(nn, 1, bB);
        TL_STAMP(18);
        sum += epi_last(net, 2);
        TL_STAMP(19);
      }
      if (!hf) finalize(tile, sum, cur_active, cur_cd, cur_ll);
      tile = next_tile;
      have = have_next;
      cur_active = hold_active; cur_cd = hold_cd; cur_ll = hold_ll;
    }
  } else {
    for (int64_t tile = tile0; tile < n_tiles; tile += tile_step) {
      const bool active = nx_mask != 0;
      const uint32_t cd_in = nx_cd;
      const double ll = nx_ll;
      stage_a0(tile, active);
      tmem_wait_st();
      fetch(tile + tile_step);
      tc_fence_before();
      group_sync(g);
      float sum = 0.f;
      for (int net = 0; net < h.n_net; ++net) {
        for (int l = 0; l < h.n_hid; ++l) {
          issue(net, l, bA);
          mbar_wait(bA, phA); phA ^= 1u;
          if (l + 1 < h.n_hid) epi_hidden(l);
          else sum += epi_last(net, l);
        }
      }
      if (!hf) finalize(tile, sum, active, cd_in, ll);
    }
  }

  if (tail.partial) {
    long long cnt[NB200_N_CNT];
#pragma unroll
    for (int q = 0; q < NB200_N_CNT; ++q) cnt[q] = 0;
    cnt[NB200_CNT_RAW] = c_raw;
    cnt[NB200_CNT_CUBE_REJECT] = c_rej0;
    cnt[NB200_CNT_OVERLAP_REJECT] = c_rej1;
    cnt[NB200_CNT_NN_REJECT] = c_rej2;
    cnt[NB200_CNT_EXCLUDED] = c_rej3;
    cnt[NB200_CNT_IN_SHELL] = c_in;
    cnt[NB200_CNT_UPDATE] = c_upd;
    stat_block_reduce<TC_GROUPS * TC_GROUP_THREADS>(
        lse_acc, cnt, tail.partial + blockIdx.x);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem_slot), "r"(TC_GROUPS * TC_COLS_PER_GROUP)
                 : "memory");
  }
}

// whitened fp64 rows -> standardised, tf32-rounded fp32 rows padded to k0p
__global__ void k_standardise_tf32(const double* __restrict__ t_rows,
                                   const uint8_t* __restrict__ mask, int64_t n,
                                   int d, int k0p,
                                   const double* __restrict__ mean,
                                   const double* __restrict__ scale,
                                   float* __restrict__ xs32) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * k0p) return;
  const int64_t row = e / k0p;
  const int k = (int)(e - row * k0p);
  float v = 0.f;
  if (!mask || mask[row]) {
    if (k < d)
      v = (float)((t_rows[row * d + k] - __ldg(mean + k)) *
                  (1.0 / __ldg(scale + k)));
    else if (k == d)
      v = 1.0f;            // constant-one column that carries the biases
  }
  uint32_t rr;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rr) : "f"(v));
  xs32[e] = __uint_as_float(rr);
}

int run_mlp_tf32_streamed(const int32_t* hdr, const float* blob,
                          const float* xs32, const uint8_t* mask, int64_t n,
                          double* score_out, uint8_t* passf, uint8_t* code,
                          cudaStream_t st);

static int run_mlp_tf32(const TcHeader& h, const float* blob,
                        const float* xs32, const uint8_t* mask, int64_t n,
                        double* score_out, uint8_t* passf, uint8_t* code,
                        const TcTail& tail, int* grid_out, cudaStream_t st) {
  if ((h.magic >> 16) == 0) {
    // weights do not fit shared memory all at once: layer-at-a-time kernel
    NB_CHECK(tail.partial == nullptr, "streamed emulator has no fused tail");
    if (grid_out) *grid_out = 0;
    ProfScope prof(ST_MLP, st);
    return run_mlp_tf32_streamed((const int32_t*)&h, blob, xs32, mask, n,
                                 score_out, passf, code, st);
  }
  // one CTA per SM is REQUIRED (every CTA allocates all 512 TMEM columns):
  // ask for enough shared memory that two can never be co-resident
  size_t smem = (size_t)h.total_floats * 4;
  NB_CHECK(smem <= 220 * 1024, "emulator weights exceed shared memory");
  if (smem < 120 * 1024) smem = 120 * 1024;
  NB_CUDA(cudaFuncSetAttribute(k_mlp_tf32,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = (n + 127) / 128;
  const int n_groups = h.magic >> 16;
  int64_t grid = (n_tiles + n_groups - 1) / n_groups;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  if (tail.partial) NB_CHECK(grid <= STAT_MAX_BLOCKS, "too many partials");
  if (grid_out) *grid_out = (int)grid;
  k_mlp_tf32<<<(unsigned)grid, TC_GROUPS * TC_GROUP_THREADS, smem, st>>>(
      h, blob, xs32, mask, n, score_out, passf, code, tail);
  NB_LAUNCH_OK();
  return 0;
}

static int tc_header(const int32_t* meta_h, int bound, int j, TcHeader* h) {
  const Rec rec = record(meta_h, bound);
  const int32_t* nb = rec.nb(j);
  NB_CHECK(nb[10] >= 0 && nb[11] > 0,
           "this emulator has no tensor-core blob (architecture outside the "
           "NB200_MLP_TF32 envelope); use NB200_MLP_F64");
  memcpy(h, rec.r + nb[11], sizeof(*h));
  NB_CHECK((h->magic & 0xFFFF) == 0x7F32 && (h->magic >> 16) >= 0 &&
               (h->magic >> 16) <= TC_GROUPS,
           "corrupt tensor-core blob header");
  return 0;
}

// does neural bound j of `bound` run on the resident tensor-core kernel (the
// one that can carry the fused tail)?
bool mlp_tf32_resident(const int32_t* meta_h, int bound, int j) {
  const Rec rec = record(meta_h, bound);
  const int32_t* nb = rec.nb(j);
  if (nb[10] < 0 || nb[11] <= 0) return false;
  TcHeader h;
  memcpy(&h, rec.r + nb[11], sizeof(h));
  return (h.magic & 0xFFFF) == 0x7F32 && (h.magic >> 16) >= 1;
}

// whitened fp64 rows in, scores / pass flags out
int launch_mlp_tf32(const int32_t* meta_h, const double* data_d, int bound,
                    int j, const double* t_rows, const uint8_t* mask,
                    int64_t n, double* score_out, uint8_t* passf,
                    float* xs32_ws, cudaStream_t st) {
  TcHeader h;
  if (tc_header(meta_h, bound, j, &h)) return 1;
  const Rec rec = record(meta_h, bound);
  const int32_t* nb = rec.nb(j);
  const int64_t total = n * h.k0p;
  k_standardise_tf32<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      t_rows, mask, n, rec.d(), h.k0p, data_d + nb[5], data_d + nb[6],
      xs32_ws);
  NB_LAUNCH_OK();
  TcTail none;
  memset(&none, 0, sizeof(none));
  return run_mlp_tf32(h, (const float*)(data_d + nb[10]), xs32_ws, mask, n,
                      score_out, passf, nullptr, none, nullptr, st);
}

// standardised tf32 rows in (from k_front); rejects are written into `code`
// With `partial` != nullptr the likelihood and the per-block shell sums are
// fused in (one StatPartial per CTA, *n_partial_out of them).
int launch_mlp_tf32_rows(const int32_t* meta_h, const double* data_d,
                         int bound, int j, const float* xs32,
                         const uint8_t* mask, int64_t n, uint8_t* code,
                         double log_l_min, double* log_l, void* partial,
                         int* n_partial_out, cudaStream_t st) {
  TcHeader h;
  if (tc_header(meta_h, bound, j, &h)) return 1;
  const Rec rec = record(meta_h, bound);
  TcTail tail;
  memset(&tail, 0, sizeof(tail));
  tail.log_l = log_l;
  tail.partial = (StatPartial*)partial; tail.log_l_min = log_l_min;
  if ((h.magic >> 16) == 0) tail.partial = nullptr;   // streamed: no tail
  return run_mlp_tf32(h, (const float*)(data_d + rec.nb(j)[10]), xs32, mask,
                      n, nullptr, nullptr, code, tail, n_partial_out, st);
}

#ifdef NB200_TIMELINE
int debug_timeline(long long* out, int n) {
  int cnt = 0;
  cudaMemcpyFromSymbol(&cnt, g_tl_n, sizeof(int));
  if (cnt > n) cnt = n;
  cudaMemcpyFromSymbol(out, g_tl, sizeof(long long) * cnt);
  int zero = 0;
  cudaMemcpyToSymbol(g_tl_n, &zero, sizeof(int));
  return cnt;
}
#endif

}  // namespace nb200

#ifdef NB200_TIMELINE
extern "C" int nb200_debug_timeline(long long* out, int n) {
  return nb200::debug_timeline(out, n);
}
#endif
