// Layer-at-a-time tensor-core emulator for networks whose weights do not fit
// shared memory all at once (e.g. BASELINE config 3: 50 -> 4 x 128 -> 1, four
// networks = 1 MB of tf32 weights).
//
// Same arithmetic and the same TMEM-resident A operand as k_mlp_tf32
// (nb200_mlp_tc.cu), but one launch per (network, hidden layer): the layer's
// weights (<= 200 KB) are fetched once per persistent CTA by a TMA bulk copy,
// activations stream through global memory as tf32 rows
// ([n, round8(fan_out + 1)], constant-one column included, so biases still
// ride on the MMA), the last hidden layer folds the fan-out-1 output layer
// into its epilogue and accumulates the per-network score.  HBM traffic is
// ~4 (K + N) bytes per point and layer instead of zero, which is why the
// resident kernel is preferred whenever it applies.
#include "nb200_device.cuh"
#include "nb200_tc.cuh"

namespace nb200 {

struct LayerArgs {
  int kp, np, out_w;      // padded fan_in, padded fan_out, columns written out
  int n_groups;           // tile groups per CTA (2 if TMEM allows, else 1)
  int a_col, d_col;       // TMEM columns of the A operand and the accumulator
  int w_floats;           // floats fetched to shared memory
  int wout_off, bout_off; // (last layer) offsets of w_out / b_out in that slice
  int last, first_net;
  long long n;
};

__global__ void __launch_bounds__(512, 1)
k_layer_tf32(const LayerArgs A, const float* __restrict__ w,
             const float* __restrict__ in, const uint8_t* __restrict__ mask,
             float* __restrict__ out, float* __restrict__ score) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t wbar;
  __shared__ uint64_t mbar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float part[2][128];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = tid >> 8, r = tid & 127, hf = (tid >> 7) & 1;
  if (tid == 0) {
    mbar_init(&wbar, 1);
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float* wsm = (float*)smem;
  if (warp == 0) {
    asm volatile(
        "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
        ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"
                 ::: "memory");
  }
  if (tid == 32) {
    const uint32_t bytes = (uint32_t)A.w_floats * 4u;
    mbar_expect_tx(&wbar, bytes);
    uint32_t done = 0;
    while (done < bytes) {
      const uint32_t c = min(bytes - done, 65536u);
      bulk_g2s(smem + done, (const char*)w + done, c, &wbar);
      done += c;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int cols_per_group = 512 / A.n_groups;
  const uint32_t tmem_base = tmem_slot + (uint32_t)(g * cols_per_group);
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const bool issuer_warp = (warp_u & 7) == 0;
  const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_slot, 0) +
                               (uint32_t)((warp_u >> 3) * cols_per_group);
  const uint32_t wsm_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
  mbar_wait(&wbar, 0);

  uint32_t phase = 0;
  const long long n_tiles = (A.n + 127) / 128;
  const long long tile0 = g < A.n_groups
      ? (long long)blockIdx.x * A.n_groups + g : n_tiles;
  for (long long tile = tile0; tile < n_tiles;
       tile += (long long)gridDim.x * A.n_groups) {
    const long long row = tile * 128 + r;
    const bool active = row < A.n && (!mask || mask[row]);
    // input row -> TMEM (two threads per row, alternating 8-column chunks)
    {
      const uint4* src = (const uint4*)(in + row * (long long)A.kp);
      for (int c = hf * 8; c < A.kp; c += 16) {
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
        tmem_st8(tmem_base + lane_addr + (uint32_t)(A.a_col + c), v);
      }
      tmem_wait_st();
    }
    tc_fence_before();
    group_sync(g);
    if (issuer_warp) {
      tc_fence_after();
      const uint32_t idesc = idesc_tf32(A.np);
      const uint32_t d_tmem = tmem_base_u + (uint32_t)A.d_col;
      const uint32_t a_tmem = tmem_base_u + (uint32_t)A.a_col;
      const uint64_t desc0 = smem_desc(wsm_u, 128u, (uint32_t)A.kp * 32u);
      if (elect_one()) {
        const int ks = A.kp >> 3;
        for (int s = 0; s < ks; ++s)
          mma_tf32_ts(d_tmem, a_tmem + (uint32_t)(s * 8),
                      desc0 + (uint64_t)(s * 16), idesc, s > 0 ? 1u : 0u);
        mma_commit(&mbar[g]);
      }
      __syncwarp();
    }
    mbar_wait(&mbar[g], phase);
    phase ^= 1u;
    tc_fence_after();
    const uint32_t d_addr = tmem_base + lane_addr + (uint32_t)A.d_col;
    if (!A.last) {
      float* orow = out + row * (long long)A.out_w;
      for (int c = hf * 16; c < A.np; c += 32) {
        uint32_t v[16];
        tmem_ld16(d_addr + (uint32_t)c, v);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 16; ++q)
          v[q] = __float_as_uint(fmaxf(__uint_as_float(v[q]), 0.f)) + 0x1000u;
        if (active) {
#pragma unroll
          for (int q = 0; q < 16; q += 4)
            if (c + q < A.out_w)
              *reinterpret_cast<uint4*>(orow + c + q) =
                  make_uint4(v[q], v[q + 1], v[q + 2], v[q + 3]);
        }
      }
    } else {
      const float* wout = wsm + A.wout_off;
      float acc = hf ? 0.f : wsm[A.bout_off];
      for (int c = hf * 16; c < A.np; c += 32) {
        uint32_t v[16];
        tmem_ld16(d_addr + (uint32_t)c, v);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 16; ++q)
          acc = fmaf(fmaxf(__uint_as_float(v[q]), 0.f), wout[c + q], acc);
      }
      if (hf) part[g][r] = acc;
      group_sync(g);
      if (!hf && active) {
        const float y = acc + part[g][r];
        score[row] = A.first_net ? y : score[row] + y;
      }
    }
    // the next tile's input store / MMA reuse these TMEM columns
    tc_fence_before();
    group_sync(g);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 ::"r"(tmem_slot), "r"(512) : "memory");
  }
}

__global__ void k_score_finish(const float* __restrict__ score,
                               const uint8_t* __restrict__ mask, long long n,
                               float inv_nets, double thr,
                               double* __restrict__ score_out,
                               uint8_t* __restrict__ passf,
                               uint8_t* __restrict__ code) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool active = !mask || mask[i];
  if (!active) {
    if (score_out) score_out[i] = nan("");
    return;
  }
  const double s = (double)(score[i] * inv_nets);
  if (score_out) score_out[i] = s;
  if (passf && s > thr) passf[i] = 1;
  if (code && !(s > thr)) code[i] = NB200_CODE_NN_REJECT;
}

// header layout: see TcHeader in nb200_mlp_tc.cu (int32[32])
int run_mlp_tf32_streamed(const int32_t* hdr, const float* blob,
                          const float* xs32, const uint8_t* mask, int64_t n,
                          double* score_out, uint8_t* passf, uint8_t* code,
                          cudaStream_t st) {
  const int n_net = hdr[1], n_hid = hdr[2], k0p = hdr[4], net_stride = hdr[5];
  const int* np = hdr + 8;
  const int* kp = hdr + 12;
  const int* w_off = hdr + 16;
  const int w_out_off = hdr[28], b_out_off = hdr[29];
  double thr;
  { int t[2] = {hdr[30], hdr[31]}; memcpy(&thr, t, 8); }
  int maxw = k0p;
  for (int l = 0; l < n_hid; ++l) maxw = np[l] > maxw ? np[l] : maxw;
  float *bufA = nullptr, *bufB = nullptr, *score = nullptr;
  NB_CUDA(cudaMallocAsync(&bufA, sizeof(float) * (size_t)n * maxw, st));
  NB_CUDA(cudaMallocAsync(&bufB, sizeof(float) * (size_t)n * maxw, st));
  NB_CUDA(cudaMallocAsync(&score, sizeof(float) * (size_t)n, st));
  int dev = 0, sms = 0;
  NB_CUDA(cudaGetDevice(&dev));
  NB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = (n + 127) / 128;
  for (int net = 0; net < n_net; ++net) {
    const float* src = xs32;
    for (int l = 0; l < n_hid; ++l) {
      LayerArgs A;
      memset(&A, 0, sizeof(A));
      A.kp = kp[l]; A.np = np[l];
      A.last = (l == n_hid - 1);
      A.out_w = A.last ? 0 : kp[l + 1];
      A.a_col = 0;
      A.d_col = (A.kp + 31) / 32 * 32;
      const int cols = A.d_col + (A.np + 31) / 32 * 32;
      NB_CHECK(cols <= 512, "layer too wide for tensor memory");
      A.n_groups = cols <= 256 ? 2 : 1;
      A.first_net = (net == 0);
      A.n = n;
      const float* wslice = blob + (size_t)net * net_stride + w_off[l];
      A.w_floats = A.np * A.kp;
      if (A.last) {
        // w_out and b_out follow the weight slices of the network
        NB_CHECK(w_out_off >= w_off[l] + A.w_floats, "blob layout");
        A.wout_off = w_out_off - w_off[l];
        A.bout_off = b_out_off - w_off[l];
        A.w_floats = A.bout_off + 4;
      }
      size_t smem = (size_t)A.w_floats * 4;
      NB_CHECK(smem <= 220 * 1024, "layer weights exceed shared memory");
      // one CTA per SM is REQUIRED (each allocates all 512 TMEM columns)
      if (smem < 120 * 1024) smem = 120 * 1024;
      NB_CUDA(cudaFuncSetAttribute(
          k_layer_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize,
          (int)smem));
      int64_t grid = (n_tiles + A.n_groups - 1) / A.n_groups;
      if (grid > sms) grid = sms;
      if (grid < 1) grid = 1;
      float* dst = (l & 1) ? bufB : bufA;
      k_layer_tf32<<<(unsigned)grid, 512, smem, st>>>(A, wslice, src, mask,
                                                      dst, score);
      NB_LAUNCH_OK();
      src = dst;
    }
  }
  k_score_finish<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
      score, mask, n, 1.0f / (float)n_net, thr, score_out, passf, code);
  NB_LAUNCH_OK();
  NB_CUDA(cudaFreeAsync(bufA, st));
  NB_CUDA(cudaFreeAsync(bufB, st));
  NB_CUDA(cudaFreeAsync(score, st));
  return 0;
}

}  // namespace nb200
