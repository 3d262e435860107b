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

// Grouped mode (later-bound exclusion, nautilus/sampler.py:796-801): instead
// of one array of rows the kernel walks SEGMENTS, one per (later bound,
// neural bound) pair -- rows that k_excl_prep found inside that pair's
// ellipsoids, already whitened and standardised -- with that pair's weights
// and threshold; a row whose score passes marks its candidate as excluded.
// The flattened list of 128-row tiles of all segments is split evenly over
// the CTAs; a CTA reloads the weights (TMA bulk copy out of L2) only when it
// crosses into another segment.
// Debug timeline: clock64 stamps of one epilogue thread (slot 0) and of the
// MMA issuer (slot 1) of tile group 0 in CTA 0, read back with
// nb200_debug_timeline(); compiled in only with -DNB200_TIMELINE.  The index
// lives in a register so that a stamp is one fire-and-forget store.
#ifdef NB200_TIMELINE
__device__ long long g_tl[2][1024];
#define TL_DECL int tl_i = 0; const int tl_slot = threadIdx.x == 0 ? 0 : \
    (threadIdx.x == TC_GROUPS * TC_GROUP_THREADS ? 1 : -1)
#define TL_STAMP(tag)                                                   \
  do {                                                                  \
    if (blockIdx.x == 0 && tl_slot >= 0 && tl_i < 1022) {               \
      g_tl[tl_slot][tl_i++] = (long long)(tag);                         \
      g_tl[tl_slot][tl_i++] = clock64();                                \
    }                                                                   \
  } while (0)
#else
#define TL_DECL do {} while (0)
#define TL_STAMP(tag) do {} while (0)
#endif

constexpr int TC_EPI_THREADS = TC_GROUPS * TC_GROUP_THREADS;   // 512
constexpr int TC_THREADS = TC_EPI_THREADS + 32 * TC_GROUPS;    // + issuers

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// the two threads that share a TMEM lane (warps w and w+4 of a group)
__device__ __forceinline__ void pair_sync(int id) {
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

// ---------------------------------------------------------------------------
// Warp-specialised: per tile group 8 epilogue warps (two threads per TMEM
// lane) and ONE issuer warp.  They follow the same static schedule and meet
// only through mbarriers: the issuer commits every layer's MMAs to bA/bB/bC,
// the epilogue warps arrive on s0/s1/s2/sA when the operand they produced (or
// the region they finished reading) is ready -- nobody waits at a CTA
// barrier, a fast warp runs ahead.
// F16: the kind::f16 form (fp16 operands, fp32 accumulate; header and blob of
// _pack.py:pack_tc16).  Two fp16 values share one 32-bit TMEM column of an A
// operand, so an epilogue thread packs PAIRS of accumulator columns: the two
// threads of a TMEM lane own disjoint column ranges [0, h0) and [h0, N) of
// the accumulator (h0 = split16(N)) and write their packed halves in place at
// the start of their own range -- the A operand of the next layer therefore
// sits in two pieces, which costs nothing because the issuer passes the A
// address of every K step explicitly.
__host__ __device__ __forceinline__ int split16(int n) {
  return (n + 31) / 32 * 16;
}

template <bool GROUPED, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_mlp_tf32(const TcHeader h, const float* __restrict__ blob_arg,
           const float* __restrict__ xs32_arg,
           const uint8_t* __restrict__ mask, int64_t n_arg,
           double* __restrict__ score_out, uint8_t* __restrict__ passf,
           uint8_t* __restrict__ code, const TcTail tail,
           const TcGroupArgs G) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t wbar;
  __shared__ uint64_t mbar[TC_GROUPS * 3];    // MMA completion: bA, bB, bC
  __shared__ uint64_t sbar[TC_GROUPS * 4];    // epilogue -> issuer: s0 s1 s2 sA
  __shared__ uint32_t tmem_slot;
  __shared__ float part[TC_GROUPS][128];
  __shared__ uint4 mma_tab[TC_MAX_HID];       // {d_col, a_col, idesc, k steps}
  __shared__ int a_split[TC_MAX_HID];         // F16: K steps in A's 1st piece
  __shared__ uint64_t desc_tab[TC_MAX_HID];   // weight descriptor of network 0

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const bool is_issuer = tid >= TC_EPI_THREADS;
  // tile group: 8 epilogue warps each, then one issuer warp each
  const int g = is_issuer ? (tid - TC_EPI_THREADS) >> 5 : tid >> 8;
  const int r = tid & 127;           // row in tile == TMEM lane
  const int hf = (tid >> 7) & 1;     // which half of the columns this thread
                                     // handles (warps w and w+4 share lanes)
  TL_DECL;

  double thr = __hiloint2double(h.thr_hi, h.thr_lo);
  // architectures that need more than 256 TMEM columns run ONE tile group per
  // CTA (group 1 idles); otherwise two groups interleave
  const int n_groups = h.magic >> 16;
  if (tid == 0) {
    mbar_init(&wbar, 1);
    for (int q = 0; q < TC_GROUPS * 3; ++q) mbar_init(&mbar[q], 1);
    // one arrival per epilogue warp
    for (int q = 0; q < TC_GROUPS * 4; ++q) mbar_init(&sbar[q], 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // MMA issue table: everything tcgen05.mma needs per layer, computed once
  if (tid >= 64 && tid < 64 + h.n_hid) {
    const int l = tid - 64;
    mma_tab[l] = make_uint4(
        (uint32_t)h.d_col[l],
        (uint32_t)(l == 0 ? h.a0_col : h.d_col[l - 1]),
        F16 ? idesc_f16(h.np[l]) : idesc_tf32(h.np[l]),
        (uint32_t)(h.kp[l] >> (F16 ? 4 : 3)));
    // (an 8-row group of the K-major operand is kp * 16 B of fp16 / kp * 32 B
    // of tf32 away from the next)
    desc_tab[l] = smem_desc(smem_u32(smem) + 4u * (uint32_t)h.w_off[l], 128u,
                            (uint32_t)h.kp[l] * (F16 ? 16u : 32u));
    a_split[l] = l == 0 ? (1 << 20) : split16(h.np[l - 1]) >> 4;
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot + (uint32_t)(g * TC_COLS_PER_GROUP);
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;

  uint64_t* bA = &mbar[g * 3 + 0];
  uint64_t* bB = &mbar[g * 3 + 1];
  uint64_t* bC = &mbar[g * 3 + 2];
  uint64_t* s0 = &sbar[g * 4 + 0];   // epilogue 0 done: A of layer 1 ready
  uint64_t* s1 = &sbar[g * 4 + 1];   // epilogue 1 done: A of layer 2 ready
  uint64_t* s2 = &sbar[g * 4 + 2];   // last epilogue done: D2 region free
  uint64_t* sA = &sbar[g * 4 + 3];   // input rows of a tile staged

  // ---- the rows this CTA works on ------------------------------------------
  // plain mode: ONE segment = the whole array, tiles interleaved over the
  // CTAs.  grouped mode: this CTA's share [my_lo, my_hi) of the flattened
  // tile list of all segments, walked segment by segment.
  const float* blob = blob_arg;
  const float* xs32 = xs32_arg;
  int64_t n = n_arg;
  int64_t n_tiles = (n + 127) / 128;
  int64_t tile_step = (int64_t)gridDim.x * n_groups;
  int64_t tile0 = g < n_groups ? (int64_t)blockIdx.x * n_groups + g : n_tiles;
  int64_t seg_row0 = 0;              // first row of the segment in cid[]
  long long my_lo = 0, my_hi = 0, seg_first = 0;
  int seg = 0, seg_end = 1;
  if (GROUPED) {
    seg_end = 0;
    if ((unsigned long long)G.chunk_lo < *G.n_cand) {
      long long total = 0;
      for (int p = 0; p < G.n_pairs; ++p)
        total += ((long long)G.seg_count[p] + 127) >> 7;
      my_lo = total * (long long)blockIdx.x / (long long)gridDim.x;
      my_hi = total * (long long)(blockIdx.x + 1) / (long long)gridDim.x;
      seg_end = my_hi > my_lo ? G.n_pairs : 0;
    }
  }
  uint32_t wphase = 0;
  bool loaded = false;

  // phases of the hand-over barriers run on across segments
  uint32_t p0 = 0, p1 = 0, p2 = 0, pA = 0;         // issuer side
  uint32_t phA = 0, phB = 0, phC = 0;             // epilogue side

  Lse lse_acc;
  lse_acc.init();
  int c_rej0 = 0, c_rej1 = 0, c_rej2 = 0, c_rej3 = 0, c_in = 0, c_upd = 0,
      c_raw = 0;

  for (; seg < seg_end; ++seg) {
  if (GROUPED) {
    // this CTA's tiles [t_lo, t_hi) of segment `seg` (warp-uniform)
    const long long tiles_p = ((long long)G.seg_count[seg] + 127) >> 7;
    const long long lo = my_lo > seg_first ? my_lo - seg_first : 0;
    const long long hi = my_hi - seg_first < tiles_p ? my_hi - seg_first
                                                     : tiles_p;
    seg_first += tiles_p;
    if (hi <= lo) continue;
    const PairRec pr = G.pairs[seg];
    blob = reinterpret_cast<const float*>(G.data + pr.blob_off);
    thr = G.data[pr.thr_off];
    seg_row0 = (int64_t)seg * G.seg_stride;
    xs32 = reinterpret_cast<const float*>(
        reinterpret_cast<const uint32_t*>(xs32_arg) +
        seg_row0 * (int64_t)(F16 ? h.k0p >> 1 : h.k0p));
    n = (int64_t)G.seg_count[seg];
    n_tiles = hi;
    tile_step = n_groups;
    tile0 = g < n_groups ? lo + g : n_tiles;
  }
  // ---- weights of this segment -> shared memory (TMA bulk copy) ------------
  if (loaded) {
    // every MMA that read the old weights has been waited for by the epilogue
    // warps; everyone is done with the old segment
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
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
  mbar_wait(&wbar, wphase);
  wphase ^= 1u;
  loaded = true;
  // tiles this group processes
  const int64_t my_tiles =
      tile0 < n_tiles ? (n_tiles - tile0 + tile_step - 1) / tile_step : 0;

  if (is_issuer) {
    // =====================================================================
    // MMA issuer: one elected lane issues, the warp stays converged
    // =====================================================================
    const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t net_step16 = (uint32_t)h.net_stride >> 2;
    auto issue = [&](int net, int l, uint64_t* bar) {
      tc_fence_after();
      if (elect_one()) {
        const uint4 t = mma_tab[l];
        // descriptor address field counts 16-byte units
        const uint64_t desc0 =
            desc_tab[l] + (uint64_t)((uint32_t)net * net_step16);
        const uint32_t d_tmem = tmem_base_u + t.x;
        const uint32_t a_tmem = tmem_base_u + t.y;
        const int ks = (int)t.w;
        if (F16) {
          // A's second piece starts at column h0 of the previous
          // accumulator: K step s >= sp lives 8 (s - sp) columns behind it
          const int sp = a_split[l];
          for (int s = 0; s < ks; ++s) {
            const uint32_t a_col = s < sp ? (uint32_t)(s * 8)
                                          : (uint32_t)(sp * 16 + (s - sp) * 8);
            mma_f16_ts(d_tmem, a_tmem + a_col, desc0 + (uint64_t)(s * 16),
                       t.z, s > 0 ? 1u : 0u);
          }
        } else {
          for (int s = 0; s < ks; ++s) {
            // one K-step = 2 core matrices = 256 B -> +16 in the 16-B units
            // of the descriptor's start-address field
            mma_tf32_ts(d_tmem, a_tmem + (uint32_t)(s * 8),
                        desc0 + (uint64_t)(s * 16), t.z, s > 0 ? 1u : 0u);
          }
        }
        mma_commit(bar);
      }
      __syncwarp();
    };
    const int N = h.n_net;
    if (h.n_hid == 3) {
      // instance i = (tile, network); MMAs execute in issue order, so region
      // reuse between MMAs needs no wait -- only operands written by the
      // epilogue warps do:
      //   L1(i)   after epilogue0(i);  L0(i+1) right behind it (after the
      //   next tile's rows are staged if i+1 opens a tile);
      //   L2(i)   after epilogue1(i) and after the last epilogue of i-1 has
      //   read the D2 region.
      const int64_t M = my_tiles * N;
      if (M > 0) {
        mbar_wait(sA, pA); pA ^= 1u;
        issue(0, 0, bA);
      }
      int net = 0;
      for (int64_t i = 0; i < M; ++i) {
        mbar_wait(s0, p0); p0 ^= 1u;
        TL_STAMP(20);
        issue(net, 1, bB);
        const int nn = net + 1 == N ? 0 : net + 1;
        if (i + 1 < M) {
          if (nn == 0) { mbar_wait(sA, pA); pA ^= 1u; }
          issue(nn, 0, bA);
        }
        TL_STAMP(21);
        mbar_wait(s1, p1); p1 ^= 1u;
        if (i > 0) { mbar_wait(s2, p2); p2 ^= 1u; }
        TL_STAMP(22);
        issue(net, 2, bC);
        TL_STAMP(23);
        net = nn;
      }
    } else {
      // serial schedule: one layer at a time, s0 carries every hand-over
      for (int64_t t = 0; t < my_tiles; ++t)
        for (int net = 0; net < N; ++net)
          for (int l = 0; l < h.n_hid; ++l) {
            mbar_wait(s0, p0); p0 ^= 1u;
            issue(net, l, bA);
          }
    }
  } else {
    // =====================================================================
    // epilogue warps
    // =====================================================================
    const bool with_tail = tail.partial != nullptr;
    // The inputs of the NEXT tile are fetched into registers while the
    // current tile runs: this half's 16 columns of the standardised row
    // (k0p <= 32: 4 x 16 B), the candidate flag, and for the fused tail the
    // disposition byte and the likelihood the front kernel already
    // evaluated.  None of the loads depends on another one.
    // 32-bit words of one input row: k0p tf32 values, or k0p / 2 fp16 pairs
    const int xw = F16 ? h.k0p >> 1 : h.k0p;
    const uint32_t* xrows = reinterpret_cast<const uint32_t*>(xs32);
    const bool prefetch = xw <= 32;
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
          const uint4* src = (const uint4*)(xrows + rw * (int64_t)xw) + hf * 4;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (hf * 16 + q * 4 < xw) pre[q] = __ldg(src + q);
        }
      }
    };
    // this warp's part of a hand-over to the issuer: TMEM stores complete,
    // ordered before the arrival
    auto signal = [&](uint64_t* bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // standardised input rows of tile t -> TMEM (A operand of layer 0)
    auto stage_a0 = [&](int64_t t, bool active) {
      if (prefetch) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c = hf * 16 + q * 8;
          if (c < xw) {
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
        const uint4* src = (const uint4*)(xrows + rw * (int64_t)xw);
        for (int c = hf * 8; c < xw; c += 16) {
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
      tmem_wait_st();
      signal(sA);
    };
    // ReLU + tf32 rounding of layer l's accumulator, written back in place
    // as the next layer's A operand (the bias was added by the MMA through
    // the constant-one column, see _pack.py:pack_tc)
    auto epi_hidden = [&](int l, uint64_t* done) {
      tc_fence_after();
      const uint32_t d_addr = tmem_base + lane_addr + (uint32_t)h.d_col[l];
      if (F16) {
        // this thread's own accumulator columns [c0, c1): ReLU, round to
        // fp16, two values per 32-bit column, written back at the start of
        // the range (always behind what has been read)
        const int h0 = split16(h.np[l]);
        const int c0 = hf ? h0 : 0, c1 = hf ? h.np[l] : h0;
        int c = c0;
        for (; c + 32 <= c1; c += 32) {
          uint32_t v[32], w[16];
          tmem_ld32(d_addr + (uint32_t)c, v);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 16; ++q) w[q] = relu_f16x2(v[2 * q], v[2 * q + 1]);
          tmem_st16(d_addr + (uint32_t)(c0 + ((c - c0) >> 1)), w);
        }
        if (c < c1) {                      // a last group of 16 columns
          uint32_t v[16], w[8];
          tmem_ld16(d_addr + (uint32_t)c, v);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 8; ++q) w[q] = relu_f16x2(v[2 * q], v[2 * q + 1]);
          tmem_st8(d_addr + (uint32_t)(c0 + ((c - c0) >> 1)), w);
        }
        tmem_wait_st();
        signal(done);
        return;
      }
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
      signal(done);
    };
    // last hidden layer, the fan_out-1 output layer folded in registers;
    // `done` (may be null) tells the issuer the region has been read
    auto epi_last = [&](int net, int l, uint64_t* done) -> float {
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
      if (done) signal(done);
      // the partner half's partial dot product: two pair barriers, the
      // second one frees `part` for the next network
      const int pid = 1 + g * 4 + (warp & 3);
      if (hf) part[g][r] = acc;
      pair_sync(pid);
      const float tot = hf ? 0.f : acc + part[g][r];
      pair_sync(pid);
      return tot;
    };
    // half 0: score -> decision -> outputs of one row, and with the fused
    // tail the disposition histogram and the log-sum-exp of the likelihood
    // that k_front evaluated for this row
    auto finalize = [&](int64_t tile, float sum, bool active, uint32_t cd_in,
                        double ll) {
      const int64_t row = tile * 128 + r;
      bool accepted = false;
      if (GROUPED) {
        // inside this later bound: the candidate leaves the shell
        if (active && (double)(sum / (float)h.n_net) > thr)
          G.excl[G.cid[seg_row0 + row]] = 1;
        return;
      }
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
        } else if (active) {
          tail.log_l[row] = nan("");      // the front's value is void now
          c_rej2 += 1;
        } else {
          c_rej0 += cd_in == NB200_CODE_CUBE_REJECT;
          c_rej1 += cd_in == NB200_CODE_OVERLAP_REJECT;
          c_rej2 += cd_in == NB200_CODE_NN_REJECT;
          c_rej3 += cd_in == NB200_CODE_EXCLUDED;
        }
      }
    };

    const int N = h.n_net;
    fetch(tile0);
    if (h.n_hid == 3) {
      // One continuous stream of (tile, network) instances per group, two in
      // flight: while these warps run one instance's epilogue the tensor
      // pipe runs the other's next layer, and the stream does not drain at
      // tile boundaries.
      int64_t tile = tile0;
      bool have = tile < n_tiles;
      bool cur_active = false, hold_active = false;
      uint32_t cur_cd = NB200_CODE_IN_SHELL, hold_cd = NB200_CODE_IN_SHELL;
      double cur_ll = 0.0, hold_ll = 0.0;
      if (have) {
        TL_STAMP(1);
        cur_active = nx_mask != 0; cur_cd = nx_cd; cur_ll = nx_ll;
        stage_a0(tile, cur_active);
        fetch(tile + tile_step);
        mbar_wait(bA, phA); phA ^= 1u;
        epi_hidden(0, s0);
      }
      while (have) {
        const int64_t next_tile = tile + tile_step;
        const bool have_next = next_tile < n_tiles;
        float sum = 0.f;
        for (int net = 0; net < N; ++net) {
          const bool last = net + 1 == N;
          const bool more = !last || have_next;
          if (last && have_next) {
            // every network of this tile has read the input rows (the last
            // L0 was awaited before its epilogue): stage the next tile's
            hold_active = nx_mask != 0; hold_cd = nx_cd; hold_ll = nx_ll;
            stage_a0(next_tile, hold_active);
            fetch(next_tile + tile_step);
            TL_STAMP(1);
          }
          TL_STAMP(10);
          mbar_wait(bB, phB); phB ^= 1u;       // D1(i) ready
          TL_STAMP(11);
          epi_hidden(1, s1);
          TL_STAMP(13);
          if (more) {
            mbar_wait(bA, phA); phA ^= 1u;     // D0(i+1) ready
            TL_STAMP(15);
            epi_hidden(0, s0);
            TL_STAMP(16);
          }
          mbar_wait(bC, phC); phC ^= 1u;       // D2(i) ready
          TL_STAMP(17);
          sum += epi_last(net, 2, more ? s2 : nullptr);
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
        // (for the tiles after the first this arrival also tells the issuer
        // that the last epilogue of the previous tile is over)
        stage_a0(tile, active);
        if (lane == 0) mbar_arrive(s0);       // s0 carries every hand-over
        fetch(tile + tile_step);
        float sum = 0.f;
        for (int net = 0; net < N; ++net) {
          for (int l = 0; l < h.n_hid; ++l) {
            mbar_wait(bA, phA); phA ^= 1u;
            if (l + 1 < h.n_hid) epi_hidden(l, s0);
            else sum += epi_last(net, l, net + 1 < N ? s0 : nullptr);
          }
        }
        if (!hf) finalize(tile, sum, active, cd_in, ll);
      }
    }
  }

  }  // segments

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
    stat_block_reduce<TC_THREADS>(lse_acc, cnt, tail.partial + blockIdx.x);
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

// ... and the fp16 form: k0p halves per row, two to a 32-bit word
__global__ void k_standardise_f16(const double* __restrict__ t_rows,
                                  const uint8_t* __restrict__ mask, int64_t n,
                                  int d, int k0p,
                                  const double* __restrict__ mean,
                                  const double* __restrict__ scale,
                                  uint32_t* __restrict__ xs16) {
  const int words = k0p >> 1;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * words) return;
  const int64_t row = e / words;
  const int k = 2 * (int)(e - row * words);
  float v[2] = {0.f, 0.f};
  if (!mask || mask[row]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (k + q < d)
        v[q] = (float)((t_rows[row * d + k + q] - __ldg(mean + k + q)) *
                       (1.0 / __ldg(scale + k + q)));
      else if (k + q == d)
        v[q] = 1.0f;
    }
  }
  xs16[e] = pack_f16x2(v[0], v[1]);
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
  NB_CUDA(cudaFuncSetAttribute(k_mlp_tf32<false, false>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem));
  NB_CUDA(cudaFuncSetAttribute(k_mlp_tf32<false, true>,
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
  TcGroupArgs none;
  memset(&none, 0, sizeof(none));
  if ((h.magic & 0xFFFF) == 0x7F16)
    k_mlp_tf32<false, true><<<(unsigned)grid, TC_THREADS, smem, st>>>(
        h, blob, xs32, mask, n, score_out, passf, code, tail, none);
  else
    k_mlp_tf32<false, false><<<(unsigned)grid, TC_THREADS, smem, st>>>(
        h, blob, xs32, mask, n, score_out, passf, code, tail, none);
  NB_LAUNCH_OK();
  return 0;
}

// Grouped mode: every (later bound, neural bound) segment of one candidate
// chunk in ONE launch.  `h` is the architecture shared by all the pairs
// (checked by the caller); thresholds and weights come from G.pairs.
int run_mlp_tf32_grouped(const int32_t* hdr32, const float* xs_segments,
                         const TcGroupArgs& G, cudaStream_t st) {
  TcHeader h;
  memcpy(&h, hdr32, sizeof(h));
  NB_CHECK((h.magic >> 16) >= 1, "grouped emulator needs resident weights");
  size_t smem = (size_t)h.total_floats * 4;
  NB_CHECK(smem <= 220 * 1024, "emulator weights exceed shared memory");
  if (smem < 120 * 1024) smem = 120 * 1024;
  const bool f16 = (h.magic & 0xFFFF) == 0x7F16;
  if (f16)
    NB_CUDA(cudaFuncSetAttribute(k_mlp_tf32<true, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  else
    NB_CUDA(cudaFuncSetAttribute(k_mlp_tf32<true, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TcTail tail;
  memset(&tail, 0, sizeof(tail));
  if (f16)
    k_mlp_tf32<true, true><<<(unsigned)sms, TC_THREADS, smem, st>>>(
        h, nullptr, xs_segments, nullptr, 0, nullptr, nullptr, nullptr, tail,
        G);
  else
    k_mlp_tf32<true, false><<<(unsigned)sms, TC_THREADS, smem, st>>>(
        h, nullptr, xs_segments, nullptr, 0, nullptr, nullptr, nullptr, tail,
        G);
  NB_LAUNCH_OK();
  return 0;
}

// header and weight blob of neural bound j for the arithmetic `mode`
// (NB200_MLP_TF32: pack_tc, NB200_MLP_F16: pack_tc16, stored right behind)
static int tc_header(const int32_t* meta_h, const double* data_d, int bound,
                     int j, int mode, TcHeader* h, const float** blob) {
  const Rec rec = record(meta_h, bound);
  const int32_t* nb = rec.nb(j);
  NB_CHECK(nb[10] >= 0 && nb[11] > 0,
           "this emulator has no tensor-core blob (architecture outside the "
           "NB200_MLP_TF32 envelope); use NB200_MLP_F64");
  if (mode == NB200_MLP_F16) {
    memcpy(h, rec.r + nb[11] + TC_HDR_WORDS, sizeof(*h));
    NB_CHECK((h->magic & 0xFFFF) == 0x7F16 && (h->magic >> 16) >= 1 &&
                 (h->magic >> 16) <= TC_GROUPS,
             "this emulator has no fp16 tensor-core blob (architecture "
             "outside the NB200_MLP_F16 envelope); use NB200_MLP_TF32");
    if (blob) *blob = (const float*)(data_d + h->b_off[0]);
    return 0;
  }
  memcpy(h, rec.r + nb[11], sizeof(*h));
  NB_CHECK((h->magic & 0xFFFF) == 0x7F32 && (h->magic >> 16) >= 0 &&
               (h->magic >> 16) <= TC_GROUPS,
           "corrupt tensor-core blob header");
  if (blob) *blob = (const float*)(data_d + nb[10]);
  return 0;
}

// the tensor-core header the grouped exclusion compares / launches with
const int32_t* tc_header_words(const int32_t* meta_h, int bound, int j,
                               int mode) {
  const Rec rec = record(meta_h, bound);
  const int32_t* nb = rec.nb(j);
  if (nb[3] <= 0 || nb[10] < 0 || nb[11] <= 0) return nullptr;
  const int32_t* h = rec.r + nb[11] + (mode == NB200_MLP_F16 ? TC_HDR_WORDS
                                                             : 0);
  const int magic = mode == NB200_MLP_F16 ? 0x7F16 : 0x7F32;
  if ((h[0] & 0xFFFF) != magic || (h[0] >> 16) < 1) return nullptr;
  return h;
}

// does neural bound j of `bound` run on the resident tensor-core kernel (the
// one that can carry the fused tail)?
bool mlp_tf32_resident(const int32_t* meta_h, int bound, int j, int mode) {
  return tc_header_words(meta_h, bound, j, mode) != nullptr;
}

// whitened fp64 rows in, scores / pass flags out
int launch_mlp_tf32(const int32_t* meta_h, const double* data_d, int bound,
                    int j, const double* t_rows, const uint8_t* mask,
                    int64_t n, double* score_out, uint8_t* passf,
                    float* xs32_ws, int mode, cudaStream_t st) {
  TcHeader h;
  const float* blob = nullptr;
  if (tc_header(meta_h, data_d, bound, j, mode, &h, &blob)) return 1;
  const Rec rec = record(meta_h, bound);
  const int32_t* nb = rec.nb(j);
  if (mode == NB200_MLP_F16) {
    const int64_t total = n * (h.k0p >> 1);
    k_standardise_f16<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        t_rows, mask, n, rec.d(), h.k0p, data_d + nb[5], data_d + nb[6],
        (uint32_t*)xs32_ws);
  } else {
    const int64_t total = n * h.k0p;
    k_standardise_tf32<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        t_rows, mask, n, rec.d(), h.k0p, data_d + nb[5], data_d + nb[6],
        xs32_ws);
  }
  NB_LAUNCH_OK();
  TcTail none;
  memset(&none, 0, sizeof(none));
  return run_mlp_tf32(h, blob, xs32_ws, mask, n, score_out, passf, nullptr,
                      none, nullptr, st);
}

// standardised tf32 rows in (from k_front); rejects are written into `code`
// With `partial` != nullptr the likelihood and the per-block shell sums are
// fused in (one StatPartial per CTA, *n_partial_out of them).
int launch_mlp_tf32_rows(const int32_t* meta_h, const double* data_d,
                         int bound, int j, const float* xs32,
                         const uint8_t* mask, int64_t n, uint8_t* code,
                         double log_l_min, double* log_l, void* partial,
                         int* n_partial_out, int mode, cudaStream_t st) {
  TcHeader h;
  const float* blob = nullptr;
  if (tc_header(meta_h, data_d, bound, j, mode, &h, &blob)) return 1;
  TcTail tail;
  memset(&tail, 0, sizeof(tail));
  tail.log_l = log_l;
  tail.partial = (StatPartial*)partial; tail.log_l_min = log_l_min;
  if ((h.magic >> 16) == 0) tail.partial = nullptr;   // streamed: no tail
  return run_mlp_tf32(h, blob, xs32, mask, n, nullptr, nullptr, code, tail,
                      n_partial_out, st);
}

#ifdef NB200_TIMELINE
int debug_timeline(long long* out, int n) {
  // out: [2][1024] (tag, clock) pairs of the epilogue thread and the issuer
  if (n < 2048) return -1;
  cudaMemcpyFromSymbol(out, g_tl, sizeof(long long) * 2048);
  cudaMemset(nullptr, 0, 0);
  long long zero[2048];
  memset(zero, 0, sizeof(zero));
  cudaMemcpyToSymbol(g_tl, zero, sizeof(zero));
  return 2048;
}
#endif

}  // namespace nb200

#ifdef NB200_TIMELINE
extern "C" int nb200_debug_timeline(long long* out, int n) {
  return nb200::debug_timeline(out, n);
}
#endif
