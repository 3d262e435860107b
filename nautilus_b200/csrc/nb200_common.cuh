// Shared helpers of the nautilus_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nautilus_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "nautilus_b200 targets sm_100a (B200) only"
#endif

namespace nb200 {

// ---- error state ---------------------------------------------------------
extern thread_local char g_err[512];
extern int64_t g_launches;

inline int fail(const char* fmt, const char* a = "", long long b = 0,
                long long c = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b, c);
  return 1;
}

#define NB_CHECK(cond, msg)                                              \
  do {                                                                   \
    if (!(cond)) return nb200::fail("nautilus_b200: %s", msg);          \
  } while (0)

#define NB_CUDA(expr)                                                        \
  do {                                                                       \
    cudaError_t e_ = (expr);                                                 \
    if (e_ != cudaSuccess) {                                                 \
      snprintf(nb200::g_err, sizeof(nb200::g_err),                           \
               "nautilus_b200: CUDA error '%s' at %s:%d", cudaGetErrorString(e_), \
               __FILE__, __LINE__);                                          \
      return 2;                                                              \
    }                                                                        \
  } while (0)

#define NB_LAUNCH_OK()                 \
  do {                                 \
    nb200::g_launches++;               \
    NB_CUDA(cudaGetLastError());       \
  } while (0)

// ---- optional per-stage CUDA-event profiler (bench.py's roofline leg) -----
enum Stage {
  ST_PROPOSE = 0, ST_UNION, ST_PREP, ST_MLP, ST_GLUE, ST_LOGLIKE, ST_STATS,
  ST_COMPACT, ST_FUSED, ST_FIT, N_STAGES
};
extern bool g_prof_on;
void prof_push(int stage, cudaStream_t st, bool begin);
struct ProfScope {     // records an event pair around the launches in scope
  int stage; cudaStream_t st;
  ProfScope(int s, cudaStream_t t) : stage(s), st(t) {
    if (g_prof_on) prof_push(stage, st, true);
  }
  ~ProfScope() { if (g_prof_on) prof_push(stage, st, false); }
};

// ---- blob accessors (layout: include/nautilus_b200.h) --------------------
constexpr int HDR = 16;
constexpr int MIX_REC = 8;
constexpr int NB_REC = 12;

struct Rec {          // header of one bound record, host or device
  const int32_t* r;   // record start
  __host__ __device__ int kind() const { return r[1]; }
  __host__ __device__ int d() const { return r[2]; }
  __host__ __device__ int K() const { return r[3]; }
  __host__ __device__ int J() const { return r[4]; }
  __host__ __device__ int unit() const { return r[5]; }
  __host__ __device__ int off_cdf() const { return r[6]; }
  __host__ __device__ int max_width() const { return r[9]; }
  __host__ __device__ const int32_t* mix(int k) const {
    return r + r[7] + k * MIX_REC;
  }
  __host__ __device__ const int32_t* nb(int j) const {
    return r + r[8] + j * NB_REC;
  }
};

__host__ __device__ inline Rec record(const int32_t* meta, int bound) {
  return Rec{meta + meta[1 + bound]};
}

// ---- grouped later-bound exclusion (csrc/nb200_exclude.cu, nb200_mlp_tc.cu) --
struct PairRec { int rec_off, j, blob_off, thr_off; };   // data offsets: doubles
struct TcGroupArgs {
  int n_pairs;
  long long chunk_lo;                    // first candidate of this chunk
  long long seg_stride;                  // rows reserved per segment
  const PairRec* pairs;                  // [n_pairs]
  const unsigned int* seg_count;         // [n_pairs] rows appended
  const unsigned long long* n_cand;      // candidates in total
  const unsigned int* cid;               // [n_pairs][seg_stride] candidate ids
  const double* data;                    // the stack's parameter array
  uint8_t* excl;                         // [n_cand] out
};


// emulator arithmetic on the tensor cores?
inline bool mode_tc(int mode) {
  return mode == NB200_MLP_TF32 || mode == NB200_MLP_F16;
}

// ---- small device utilities ----------------------------------------------
__device__ __forceinline__ int row_stride(int d) { return d | 1; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace nb200
