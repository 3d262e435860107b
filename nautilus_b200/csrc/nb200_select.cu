// k-th largest of a float64 vector on the device: the live-set threshold of
// the sampler (nautilus/sampler.py:1007-1009 `log_l[-n_live]` after a full
// argsort, :1160-1164 `argsort(log_l)[-n_live:]` on every exploration step).
// MSB-first radix select on the order-preserving 64-bit image of the doubles:
// four passes of 16 bits, each a histogram over the elements that still match
// the prefix (global atomics into 65 536 bins) followed by a one-block walk
// down the bins.  O(n) reads per pass, no sort, nothing leaves the device.
#include "nb200_common.cuh"

namespace nb200 {

constexpr int SEL_BINS = 1 << 16;

struct SelState {            // lives in the workspace
  unsigned long long prefix; // high digits of the k-th largest key so far
  long long remaining;       // rank still to find inside the prefix
  long long greater;         // elements strictly greater than the answer
};

// doubles -> keys with the same order (NaN sorts above +inf; callers pass
// finite or -inf log-likelihoods)
__device__ __forceinline__ unsigned long long sel_key(double v) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double sel_value(unsigned long long k) {
  const unsigned long long u =
      (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double((long long)u);
}

__global__ void k_sel_init(SelState* st, long long k, unsigned int* hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < SEL_BINS) hist[i] = 0;
  if (i == 0) { st->prefix = 0; st->remaining = k; st->greater = 0; }
}

__global__ void k_sel_hist(const double* __restrict__ v,
                           const uint8_t* __restrict__ mask, long long n,
                           int pass, const SelState* __restrict__ st,
                           unsigned int* __restrict__ hist) {
  const int shift = 48 - 16 * pass;
  const unsigned long long prefix = st->prefix;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) continue;
    const unsigned long long key = sel_key(v[i]);
    // the digits above `shift + 16` must equal the prefix found so far
    if (pass > 0 && (key >> (shift + 16)) != (prefix >> (shift + 16))) continue;
    atomicAdd(hist + ((key >> shift) & 0xFFFF), 1u);
  }
}

// walk the bins from the top: the bin that holds the `remaining`-th largest
__global__ void __launch_bounds__(1024)
k_sel_pick(SelState* st, unsigned int* hist, int pass, double* thr_out,
           long long* greater_out) {
  __shared__ long long chunk_sum[1024];
  __shared__ int pick_chunk;
  __shared__ long long above_chunk;
  const int t = threadIdx.x;
  // thread t owns bins [64 t, 64 t + 64), highest bins first for t = 1023
  long long s = 0;
  for (int b = 0; b < 64; ++b) s += hist[t * 64 + b];
  chunk_sum[t] = s;
  __syncthreads();
  if (t == 0) {
    long long acc = 0;
    int c = 1023;
    for (; c > 0; --c) {
      if (acc + chunk_sum[c] >= st->remaining) break;
      acc += chunk_sum[c];
    }
    pick_chunk = c;
    above_chunk = acc;
  }
  __syncthreads();
  if (t == 0) {
    long long acc = above_chunk;
    int b = 63;
    for (; b > 0; --b) {
      const long long h = hist[pick_chunk * 64 + b];
      if (acc + h >= st->remaining) break;
      acc += h;
    }
    const int shift = 48 - 16 * pass;
    const unsigned long long digit = (unsigned long long)(pick_chunk * 64 + b);
    st->prefix |= digit << shift;
    st->remaining -= acc;
    st->greater += acc;
    if (pass == 3) {
      *thr_out = sel_value(st->prefix);
      *greater_out = st->greater;
    }
  }
  __syncthreads();
  for (int b = 0; b < 64; ++b) hist[t * 64 + b] = 0;   // ready for next pass
}

}  // namespace nb200

using namespace nb200;

extern "C" {

size_t nb200_select_workspace_bytes(void) {
  return sizeof(unsigned int) * SEL_BINS + 256;
}

int nb200_select_kth_largest(const double* values_d, const uint8_t* mask_d,
                             int64_t n, int64_t k, double* thr_d,
                             int64_t* greater_d, void* workspace_d,
                             size_t workspace_bytes, void* stream) {
  NB_CHECK(n >= 1 && k >= 1 && k <= n, "need 1 <= k <= n");
  NB_CHECK(values_d && thr_d && greater_d, "null argument");
  NB_CHECK(workspace_bytes >= nb200_select_workspace_bytes(),
           "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  SelState* state = (SelState*)workspace_d;
  unsigned int* hist = (unsigned int*)((char*)workspace_d + 256);
  k_sel_init<<<SEL_BINS / 256, 256, 0, st>>>(state, (long long)k, hist);
  NB_LAUNCH_OK();
  long long grid = (n + 255) / 256;
  if (grid > 1184) grid = 1184;
  for (int pass = 0; pass < 4; ++pass) {
    k_sel_hist<<<(unsigned)grid, 256, 0, st>>>(values_d, mask_d, n, pass, state,
                                               hist);
    NB_LAUNCH_OK();
    k_sel_pick<<<1, 1024, 0, st>>>(state, hist, pass, thr_d,
                                   (long long*)greater_d);
    NB_LAUNCH_OK();
  }
  return 0;
}

}  // extern "C"
