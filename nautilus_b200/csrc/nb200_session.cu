// Host-buffer *session*: the cycle for callers that live on the host (NumPy,
// ctypes) without paying nb200_cycle_host's per-call allocations, and with the
// device->host copy of batch i overlapped with the kernels of batch i + 1.
//
// This is the C-ABI form of the loop around Sampler.add_samples
// (nautilus/sampler.py:1093-1144): every submit is one raw batch of the hot
// path, every wait hands back what add_samples appends to its host arrays
// (in-shell points, their log_l) plus the integer counters and the
// log-sum-exp triple that update_shell_info (sampler.py:910-943) needs.
//
// Streams: `compute` runs H2D of the bound + kernels + compaction + the small
// D2H (counters, sums, n_out); `copy` runs the sized D2H of the compacted
// rows into the session's pinned buffers.  Raw buffers (all n proposals) are
// shared by the slots because the compute stream serialises the batches;
// compacted outputs are per slot.
#include <new>

#include "nb200_common.cuh"

namespace nb200 {

constexpr int SESSION_MAX_SLOTS = 4;
constexpr int SMALL_WORDS = NB200_N_LSE + NB200_N_CNT + 1;  // 8-byte words

struct Slot {
  double* out_points_d = nullptr;
  double* out_log_l_d = nullptr;
  unsigned long long* small_d = nullptr;   // lse[4] | counters[8] | n_out
  double* like_d = nullptr;
  double* points_h = nullptr;              // pinned
  double* log_l_h = nullptr;               // pinned
  unsigned long long* small_h = nullptr;   // pinned
  double* like_h = nullptr;                // pinned staging
  // index mode (nb200_session_set_returns): 16 bytes per in-shell proposal
  unsigned long long* out_index_d = nullptr;
  unsigned long long* index_h = nullptr;   // pinned
  int64_t copied = 0;                      // entries already on their way
  cudaEvent_t done = nullptr;
  cudaEvent_t copied_ev = nullptr;
  int busy = 0;
  int mode = 0;                            // returns mode of the batch in flight
};

}  // namespace nb200

struct nb200_session {
  int device = 0;
  int d = 0;
  int n_slots = 0;
  int64_t n_max = 0, cap = 0;
  int64_t n_meta = 0, n_data = 0;
  int like_cap = 0;
  cudaStream_t compute = nullptr, copy = nullptr, upload = nullptr;
  int32_t* meta_h = nullptr;   // pinned copy of the stack (launch shapes + H2D)
  double* data_h = nullptr;    // pinned
  // the device copy of the stack is double-buffered: a batch's upload runs on
  // the `upload` stream under the kernels of the batch before it
  int32_t* meta_buf[2] = {nullptr, nullptr};
  double* data_buf[2] = {nullptr, nullptr};
  cudaEvent_t buf_free[2] = {nullptr, nullptr};  // last reader of buffer done
  cudaEvent_t uploaded = nullptr;
  int cur = 0;
  int32_t* meta_d = nullptr;   // == meta_buf[cur]
  double* data_d = nullptr;    // == data_buf[cur]
  double* points_d = nullptr;  // raw batch, shared by the slots
  double* log_l_d = nullptr;
  uint8_t* code_d = nullptr;
  void* ws = nullptr;
  size_t wsb = 0;
  int returns = NB200_RETURN_ROWS;
  int64_t last_k = -1;         // in-shell count of the last finished batch
  // materialize: grow-only staging
  cudaStream_t aux = nullptr;
  unsigned long long* mat_index_d = nullptr;
  double* mat_points_d = nullptr;
  int64_t mat_cap = 0;
  nb200::Slot slot[nb200::SESSION_MAX_SLOTS];
};

namespace nb200 {

static void session_free(nb200_session* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->compute) cudaStreamSynchronize(s->compute);
  if (s->copy) cudaStreamSynchronize(s->copy);
  for (int i = 0; i < SESSION_MAX_SLOTS; ++i) {
    Slot& t = s->slot[i];
    cudaFree(t.out_points_d); cudaFree(t.out_log_l_d); cudaFree(t.small_d);
    cudaFree(t.like_d); cudaFree(t.out_index_d);
    cudaFreeHost(t.points_h); cudaFreeHost(t.log_l_h);
    cudaFreeHost(t.small_h); cudaFreeHost(t.like_h);
    cudaFreeHost(t.index_h);
    if (t.done) cudaEventDestroy(t.done);
    if (t.copied_ev) cudaEventDestroy(t.copied_ev);
  }
  if (s->aux) { cudaStreamSynchronize(s->aux); cudaStreamDestroy(s->aux); }
  cudaFree(s->mat_index_d); cudaFree(s->mat_points_d);
  cudaFreeHost(s->meta_h); cudaFreeHost(s->data_h);
  if (s->upload) cudaStreamSynchronize(s->upload);
  for (int b = 0; b < 2; ++b) {
    cudaFree(s->meta_buf[b]); cudaFree(s->data_buf[b]);
    if (s->buf_free[b]) cudaEventDestroy(s->buf_free[b]);
  }
  if (s->uploaded) cudaEventDestroy(s->uploaded);
  if (s->upload) cudaStreamDestroy(s->upload);
  cudaFree(s->points_d);
  cudaFree(s->log_l_d); cudaFree(s->code_d); cudaFree(s->ws);
  if (s->compute) cudaStreamDestroy(s->compute);
  if (s->copy) cudaStreamDestroy(s->copy);
  delete s;
}

#define NB_S(expr)                                                          \
  do {                                                                      \
    cudaError_t e_ = (expr);                                                \
    if (e_ != cudaSuccess) {                                                \
      snprintf(g_err, sizeof(g_err),                                        \
               "nautilus_b200: CUDA error '%s' at %s:%d",                   \
               cudaGetErrorString(e_), __FILE__, __LINE__);                 \
      session_free(s);                                                      \
      return 2;                                                             \
    }                                                                       \
  } while (0)

static int stack_shape_ok(const int32_t* meta_h, int64_t n_meta, int d) {
  NB_CHECK(meta_h && n_meta >= 2 && meta_h[0] >= 1, "empty bound stack");
  for (int b = 0; b < meta_h[0]; ++b) {
    NB_CHECK(meta_h[1 + b] > 0 && meta_h[1 + b] + HDR <= n_meta,
             "corrupt bound stack");
    NB_CHECK(record(meta_h, b).d() == d, "bounds of different n_dim");
  }
  return 0;
}

}  // namespace nb200

using namespace nb200;

extern "C" {

int nb200_session_create(const int32_t* meta_h, int64_t n_meta,
                         const double* data_h, int64_t n_data, int64_t n_max,
                         int64_t cap, int n_slots, int like_params_max,
                         nb200_session** out) {
  NB_CHECK(out != nullptr, "null output handle");
  *out = nullptr;
  NB_CHECK(n_max >= 1 && cap >= 1 && cap <= n_max, "bad n_max / cap");
  NB_CHECK(n_slots >= 1 && n_slots <= SESSION_MAX_SLOTS, "1 <= n_slots <= 4");
  NB_CHECK(n_meta >= 2 && n_data >= 0 && like_params_max >= 0, "bad sizes");
  NB_CHECK(meta_h && meta_h[0] >= 1, "empty bound stack");
  const int d = record(meta_h, 0).d();
  if (stack_shape_ok(meta_h, n_meta, d)) return 1;
  nb200_session* s = new (std::nothrow) nb200_session();
  NB_CHECK(s != nullptr, "out of host memory");
  NB_S(cudaGetDevice(&s->device));
  s->d = d; s->n_slots = n_slots; s->n_max = n_max; s->cap = cap;
  s->n_meta = n_meta; s->n_data = n_data; s->like_cap = like_params_max + 1;
  NB_S(cudaStreamCreateWithFlags(&s->compute, cudaStreamNonBlocking));
  NB_S(cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking));
  NB_S(cudaMallocHost(&s->meta_h, sizeof(int32_t) * n_meta));
  NB_S(cudaMallocHost(&s->data_h, sizeof(double) * (n_data + 1)));
  memcpy(s->meta_h, meta_h, sizeof(int32_t) * n_meta);
  if (n_data) memcpy(s->data_h, data_h, sizeof(double) * n_data);
  NB_S(cudaStreamCreateWithFlags(&s->upload, cudaStreamNonBlocking));
  NB_S(cudaEventCreateWithFlags(&s->uploaded, cudaEventDisableTiming));
  for (int b = 0; b < 2; ++b) {
    NB_S(cudaMalloc(&s->meta_buf[b], sizeof(int32_t) * n_meta));
    NB_S(cudaMalloc(&s->data_buf[b], sizeof(double) * (n_data + 1)));
    NB_S(cudaEventCreateWithFlags(&s->buf_free[b], cudaEventDisableTiming));
  }
  s->cur = 0;
  s->meta_d = s->meta_buf[0];
  s->data_d = s->data_buf[0];
  NB_S(cudaMalloc(&s->points_d, sizeof(double) * n_max * d));
  NB_S(cudaMalloc(&s->log_l_d, sizeof(double) * n_max));
  NB_S(cudaMalloc(&s->code_d, (size_t)n_max));
  // room for the grouped later-bound exclusion over the whole stack
  int pairs = 0;
  for (int b = 0; b < meta_h[0]; ++b)
    if (record(meta_h, b).kind() == 1) pairs += record(meta_h, b).J();
  s->wsb = nb200_cycle_workspace_bytes(n_max, d, pairs);
  NB_S(cudaMalloc(&s->ws, s->wsb));
  for (int i = 0; i < n_slots; ++i) {
    Slot& t = s->slot[i];
    // (out_points_d, n_max rows, is allocated by the first row-mode submit:
    // an index-mode session never needs it)
    NB_S(cudaMalloc(&t.out_log_l_d, sizeof(double) * n_max));
    NB_S(cudaMalloc(&t.small_d, 8 * SMALL_WORDS));
    NB_S(cudaMalloc(&t.like_d, sizeof(double) * s->like_cap));
    // (points_h likewise)
    NB_S(cudaMallocHost(&t.log_l_h, sizeof(double) * cap));
    NB_S(cudaMallocHost(&t.small_h, 8 * SMALL_WORDS));
    NB_S(cudaMallocHost(&t.like_h, sizeof(double) * s->like_cap));
    NB_S(cudaMalloc(&t.out_index_d, sizeof(unsigned long long) * n_max));
    NB_S(cudaMallocHost(&t.index_h, sizeof(unsigned long long) * cap));
    NB_S(cudaEventCreateWithFlags(&t.done, cudaEventDisableTiming));
    NB_S(cudaEventCreateWithFlags(&t.copied_ev, cudaEventDisableTiming));
  }
  NB_S(cudaStreamCreateWithFlags(&s->aux, cudaStreamNonBlocking));
  NB_S(cudaMemcpyAsync(s->meta_d, s->meta_h, sizeof(int32_t) * n_meta,
                       cudaMemcpyHostToDevice, s->compute));
  NB_S(cudaMemcpyAsync(s->data_d, s->data_h, sizeof(double) * n_data,
                       cudaMemcpyHostToDevice, s->compute));
  NB_S(cudaStreamSynchronize(s->compute));
  *out = s;
  return 0;
}

int nb200_session_destroy(nb200_session* s) {
  session_free(s);
  return 0;
}

int nb200_session_set_stack(nb200_session* s, const int32_t* meta_h,
                            int64_t n_meta, const double* data_h,
                            int64_t n_data) {
  NB_CHECK(s != nullptr, "null session");
  NB_CHECK(n_meta <= s->n_meta && n_data <= s->n_data,
           "new stack larger than the session's buffers");
  if (stack_shape_ok(meta_h, n_meta, s->d)) return 1;
  NB_CUDA(cudaSetDevice(s->device));
  // earlier batches may still read the staged copy: drain them first
  NB_CUDA(cudaStreamSynchronize(s->upload));
  NB_CUDA(cudaStreamSynchronize(s->compute));
  memcpy(s->meta_h, meta_h, sizeof(int32_t) * n_meta);
  if (n_data) memcpy(s->data_h, data_h, sizeof(double) * n_data);
  NB_CUDA(cudaMemcpyAsync(s->meta_d, s->meta_h, sizeof(int32_t) * n_meta,
                          cudaMemcpyHostToDevice, s->compute));
  NB_CUDA(cudaMemcpyAsync(s->data_d, s->data_h, sizeof(double) * n_data,
                          cudaMemcpyHostToDevice, s->compute));
  return 0;
}

int nb200_session_submit(nb200_session* s, int slot, int upload_stack,
                         int bound, int first_later, int n_later, int64_t n,
                         uint64_t seed, uint64_t offset, uint32_t stream_id,
                         int like_id, const double* like_params_h,
                         int n_like_params, double log_l_min, int mlp_mode) {
  NB_CHECK(s != nullptr, "null session");
  NB_CHECK(slot >= 0 && slot < s->n_slots, "slot out of range");
  Slot& t = s->slot[slot];
  NB_CHECK(!t.busy, "slot still holds an un-waited batch");
  NB_CHECK(n >= 0 && n <= s->n_max, "batch larger than the session's n_max");
  NB_CHECK(n_like_params >= 0 && n_like_params < s->like_cap,
           "too many likelihood parameters for this session");
  NB_CUDA(cudaSetDevice(s->device));
  cudaStream_t st = s->compute;
  // this step's inputs -- the serialised bound(s) and the likelihood
  // parameters -- go up on the `upload` stream, into the stack buffer the
  // batch in flight is NOT reading, so the copies run under its kernels
  bool uploads = false;
  if (upload_stack) {
    const int nxt = s->cur ^ 1;
    NB_CUDA(cudaStreamWaitEvent(s->upload, s->buf_free[nxt], 0));
    NB_CUDA(cudaMemcpyAsync(s->meta_buf[nxt], s->meta_h,
                            sizeof(int32_t) * s->n_meta,
                            cudaMemcpyHostToDevice, s->upload));
    NB_CUDA(cudaMemcpyAsync(s->data_buf[nxt], s->data_h,
                            sizeof(double) * s->n_data,
                            cudaMemcpyHostToDevice, s->upload));
    s->cur = nxt;
    s->meta_d = s->meta_buf[nxt];
    s->data_d = s->data_buf[nxt];
    uploads = true;
  }
  if (n_like_params > 0) {
    memcpy(t.like_h, like_params_h, sizeof(double) * n_like_params);
    NB_CUDA(cudaMemcpyAsync(t.like_d, t.like_h, sizeof(double) * n_like_params,
                            cudaMemcpyHostToDevice, s->upload));
    uploads = true;
  }
  if (uploads) {
    NB_CUDA(cudaEventRecord(s->uploaded, s->upload));
    NB_CUDA(cudaStreamWaitEvent(st, s->uploaded, 0));
  }
  double* lse_d = (double*)t.small_d;
  int64_t* cnt_d = (int64_t*)(t.small_d + NB200_N_LSE);
  int64_t* n_out_d = cnt_d + NB200_N_CNT;
  int rc = nb200_cycle(s->meta_h, s->meta_d, s->data_d, bound, first_later,
                       n_later, n, seed, offset, stream_id, like_id, t.like_d,
                       n_like_params, log_l_min, mlp_mode, s->points_d,
                       s->log_l_d, s->code_d, lse_d, cnt_d, s->ws, s->wsb, st);
  if (rc) return rc;
  NB_CUDA(cudaEventRecord(s->buf_free[s->cur], st));
  t.mode = s->returns;
  if (t.mode == NB200_RETURN_INDEX) {
    rc = nb200_compact_index(like_id >= 0 ? s->log_l_d : nullptr, s->code_d, n,
                             offset, (uint64_t*)t.out_index_d,
                             like_id >= 0 ? t.out_log_l_d : nullptr, n_out_d,
                             s->ws, s->wsb, st);
  } else {
    if (!t.out_points_d) {
      // compaction may write up to n rows before the capacity check can run
      NB_CUDA(cudaMalloc(&t.out_points_d,
                         sizeof(double) * s->n_max * s->d));
    }
    rc = nb200_compact(s->points_d, like_id >= 0 ? s->log_l_d : nullptr,
                       s->code_d, n, s->d, t.out_points_d,
                       like_id >= 0 ? t.out_log_l_d : nullptr, n_out_d, s->ws,
                       s->wsb, st);
  }
  if (rc) return rc;
  NB_CUDA(cudaMemcpyAsync(t.small_h, t.small_d, 8 * SMALL_WORDS,
                          cudaMemcpyDeviceToHost, st));
  NB_CUDA(cudaEventRecord(t.done, st));
  t.busy = like_id >= 0 ? 2 : 1;
  if (t.mode == NB200_RETURN_INDEX) {
    // the copy does not wait for the host to learn the row count: a little
    // more than the previous batch kept is sent on its way behind the
    // kernels (wait tops up the rare shortfall)
    int64_t guess = s->last_k < 0 ? s->cap : s->last_k + s->last_k / 4 + 1024;
    if (guess > s->cap) guess = s->cap;
    if (guess > n) guess = n;
    t.copied = guess;
    NB_CUDA(cudaStreamWaitEvent(s->copy, t.done, 0));
    if (guess > 0) {
      NB_CUDA(cudaMemcpyAsync(t.index_h, t.out_index_d,
                              sizeof(unsigned long long) * guess,
                              cudaMemcpyDeviceToHost, s->copy));
      if (like_id >= 0)
        NB_CUDA(cudaMemcpyAsync(t.log_l_h, t.out_log_l_d,
                                sizeof(double) * guess,
                                cudaMemcpyDeviceToHost, s->copy));
    }
    NB_CUDA(cudaEventRecord(t.copied_ev, s->copy));
  }
  return 0;
}

int nb200_session_wait(nb200_session* s, int slot, const double** points_h,
                       const double** log_l_h, int64_t* n_out, double* lse_h,
                       int64_t* counters_h) {
  NB_CHECK(s != nullptr, "null session");
  NB_CHECK(slot >= 0 && slot < s->n_slots, "slot out of range");
  Slot& t = s->slot[slot];
  NB_CHECK(t.busy, "nothing was submitted on this slot");
  NB_CHECK(t.mode == NB200_RETURN_ROWS,
           "this batch was submitted in index mode: use "
           "nb200_session_wait_index");
  NB_CUDA(cudaSetDevice(s->device));
  NB_CUDA(cudaEventSynchronize(t.done));
  const bool with_ll = t.busy == 2;
  t.busy = 0;
  const int64_t k = (int64_t)t.small_h[NB200_N_LSE + NB200_N_CNT];
  s->last_k = k;
  if (k > s->cap)
    return fail("nautilus_b200: %s (need %lld rows, cap %lld)",
                "session output capacity too small", k, s->cap);
  if (k > 0) {
    if (!t.points_h)
      NB_CUDA(cudaMallocHost(&t.points_h, sizeof(double) * s->cap * s->d));
    // the kernels of the next batch keep running on `compute` meanwhile
    NB_CUDA(cudaMemcpyAsync(t.points_h, t.out_points_d,
                            sizeof(double) * k * s->d, cudaMemcpyDeviceToHost,
                            s->copy));
    if (with_ll)
      NB_CUDA(cudaMemcpyAsync(t.log_l_h, t.out_log_l_d, sizeof(double) * k,
                              cudaMemcpyDeviceToHost, s->copy));
    NB_CUDA(cudaStreamSynchronize(s->copy));
  }
  if (points_h) *points_h = t.points_h;
  if (log_l_h) *log_l_h = with_ll ? t.log_l_h : nullptr;
  if (n_out) *n_out = k;
  if (lse_h) memcpy(lse_h, t.small_h, sizeof(double) * NB200_N_LSE);
  if (counters_h)
    memcpy(counters_h, t.small_h + NB200_N_LSE,
           sizeof(int64_t) * NB200_N_CNT);
  return 0;
}

int nb200_session_set_returns(nb200_session* s, int what) {
  NB_CHECK(s != nullptr, "null session");
  NB_CHECK(what == NB200_RETURN_ROWS || what == NB200_RETURN_INDEX,
           "returns must be NB200_RETURN_ROWS or NB200_RETURN_INDEX");
  s->returns = what;
  return 0;
}

int nb200_session_wait_index(nb200_session* s, int slot,
                             const uint64_t** index_h,
                             const double** log_l_h, int64_t* n_out,
                             double* lse_h, int64_t* counters_h) {
  NB_CHECK(s != nullptr, "null session");
  NB_CHECK(slot >= 0 && slot < s->n_slots, "slot out of range");
  Slot& t = s->slot[slot];
  NB_CHECK(t.busy, "nothing was submitted on this slot");
  NB_CHECK(t.mode == NB200_RETURN_INDEX,
           "this batch was submitted in row mode: use nb200_session_wait");
  NB_CUDA(cudaSetDevice(s->device));
  NB_CUDA(cudaEventSynchronize(t.copied_ev));
  const bool with_ll = t.busy == 2;
  t.busy = 0;
  const int64_t k = (int64_t)t.small_h[NB200_N_LSE + NB200_N_CNT];
  s->last_k = k;
  if (k > s->cap)
    return fail("nautilus_b200: %s (need %lld rows, cap %lld)",
                "session output capacity too small", k, s->cap);
  if (k > t.copied) {      // the guess fell short: fetch the rest
    const int64_t rest = k - t.copied;
    NB_CUDA(cudaMemcpyAsync(t.index_h + t.copied, t.out_index_d + t.copied,
                            sizeof(unsigned long long) * rest,
                            cudaMemcpyDeviceToHost, s->copy));
    if (with_ll)
      NB_CUDA(cudaMemcpyAsync(t.log_l_h + t.copied, t.out_log_l_d + t.copied,
                              sizeof(double) * rest, cudaMemcpyDeviceToHost,
                              s->copy));
    NB_CUDA(cudaStreamSynchronize(s->copy));
  }
  if (index_h) *index_h = (const uint64_t*)t.index_h;
  if (log_l_h) *log_l_h = with_ll ? t.log_l_h : nullptr;
  if (n_out) *n_out = k;
  if (lse_h) memcpy(lse_h, t.small_h, sizeof(double) * NB200_N_LSE);
  if (counters_h)
    memcpy(counters_h, t.small_h + NB200_N_LSE,
           sizeof(int64_t) * NB200_N_CNT);
  return 0;
}

int nb200_session_materialize(nb200_session* s, int bound, uint64_t seed,
                              uint32_t stream_id, int mlp_mode,
                              const uint64_t* index_h, int64_t k,
                              double* points_out_h) {
  NB_CHECK(s != nullptr, "null session");
  NB_CHECK(k >= 0, "negative k");
  if (k == 0) return 0;
  NB_CHECK(index_h && points_out_h, "null argument");
  NB_CUDA(cudaSetDevice(s->device));
  if (k > s->mat_cap) {
    NB_CUDA(cudaStreamSynchronize(s->aux));
    cudaFree(s->mat_index_d); cudaFree(s->mat_points_d);
    s->mat_index_d = nullptr; s->mat_points_d = nullptr; s->mat_cap = 0;
    const int64_t cap = k + k / 2;
    NB_CUDA(cudaMalloc(&s->mat_index_d, sizeof(unsigned long long) * cap));
    NB_CUDA(cudaMalloc(&s->mat_points_d, sizeof(double) * cap * s->d));
    s->mat_cap = cap;
  }
  // the stack of the batches in flight is the one to regenerate from: order
  // behind the compute stream's pending uploads
  cudaEvent_t ev;
  NB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  NB_CUDA(cudaEventRecord(ev, s->compute));
  NB_CUDA(cudaStreamWaitEvent(s->aux, ev, 0));
  NB_CUDA(cudaEventDestroy(ev));
  NB_CUDA(cudaMemcpyAsync(s->mat_index_d, index_h,
                          sizeof(unsigned long long) * k,
                          cudaMemcpyHostToDevice, s->aux));
  int rc = nb200_materialize(s->meta_h, s->meta_d, s->data_d, bound, seed,
                             stream_id, mlp_mode,
                             (const uint64_t*)s->mat_index_d, k,
                             s->mat_points_d, s->aux);
  if (rc) return rc;
  NB_CUDA(cudaMemcpyAsync(points_out_h, s->mat_points_d,
                          sizeof(double) * k * s->d, cudaMemcpyDeviceToHost,
                          s->aux));
  NB_CUDA(cudaStreamSynchronize(s->aux));
  return 0;
}

}  // extern "C"
