// fp64 kernels of the nautilus_b200 cycle and their C-ABI entry points.
//
// One thread owns one point; the point's row lives in shared memory (stride
// d|1), ellipsoid factors and MLP weights are read through the read-only path
// with warp-uniform addresses (one L1 broadcast per load).  See DESIGN.md for
// the data layout and the roofline of each kernel.
#include "nb200_device.cuh"
#include "nb200_rng.cuh"

#include <stdlib.h>
#include <vector>

namespace nb200 {

thread_local char g_err[512] = "";
int64_t g_launches = 0;

// ---- profiler ------------------------------------------------------------
bool g_prof_on = false;
struct ProfRec { cudaEvent_t a, b; int stage; };
static std::vector<ProfRec> g_prof;
void prof_push(int stage, cudaStream_t st, bool begin) {
  if (begin) {
    ProfRec r;
    r.stage = stage;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    g_prof.push_back(r);
  } else {
    for (size_t i = g_prof.size(); i-- > 0;)
      if (g_prof[i].stage == stage) { cudaEventRecord(g_prof[i].b, st); break; }
  }
}

static inline int threads_for(int d) { return d <= 48 ? 128 : 64; }
static inline size_t smem_rows(int threads, int d, int rows) {
  return (size_t)threads * (size_t)(d | 1) * sizeof(double) * rows;
}
static inline unsigned blocks_for(int64_t n, int threads) {
  return (unsigned)((n + threads - 1) / threads);
}
template <typename K>
static int opt_in_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    NB_CUDA(cudaFuncSetAttribute(
        kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  return 0;
}

// ==========================================================================
// Ellipsoid primitives (nautilus/bounds/basic.py:318-381)
// ==========================================================================

__global__ void k_ell_transform(const double* __restrict__ points, int64_t n,
                                int d, const double* __restrict__ c,
                                const double* __restrict__ M, int inverse,
                                double* __restrict__ out) {
  extern __shared__ double sm[];
  const int stride = row_stride(d);
  double* rowA = sm;
  double* rowB = sm + (size_t)blockDim.x * stride;
  const int64_t base = (int64_t)blockIdx.x * blockDim.x;
  const int nrows = (int)min((int64_t)blockDim.x, n - base);
  load_rows(points, base, nrows, d, stride, rowA);
  __syncthreads();
  if ((int)threadIdx.x < nrows) {
    double* x = rowA + threadIdx.x * stride;
    double* s = rowB + threadIdx.x * stride;
    if (!inverse) {
      for (int j = 0; j < d; ++j) s[j] = x[j] - __ldg(c + j);
      matvec_rows(M, d, 0, s, [&](int i, double v) { x[i] = v; });
    } else {
      for (int j = 0; j < d; ++j) s[j] = x[j];
      matvec_rows(M, d, 0, s,
                  [&](int i, double v) { x[i] = v + __ldg(c + i); });
    }
  }
  __syncthreads();
  store_rows(out, base, nrows, d, stride, rowA);
}

__global__ void k_ell_contains(const double* __restrict__ points, int64_t n,
                               int d, const double* __restrict__ c,
                               const double* __restrict__ Binv,
                               uint8_t* __restrict__ out,
                               double* __restrict__ r2_out) {
  extern __shared__ double sm[];
  const int stride = row_stride(d);
  double* rowA = sm;
  double* rowB = sm + (size_t)blockDim.x * stride;
  const int64_t base = (int64_t)blockIdx.x * blockDim.x;
  const int nrows = (int)min((int64_t)blockDim.x, n - base);
  load_rows(points, base, nrows, d, stride, rowA);
  __syncthreads();
  if ((int)threadIdx.x < nrows) {
    const double r2 =
        whiten_r2(rowA + threadIdx.x * stride, nullptr, d, c, Binv, 0,
                  rowB + threadIdx.x * stride, nullptr);
    out[base + threadIdx.x] = r2 < 1.0 ? 1 : 0;
    if (r2_out) r2_out[base + threadIdx.x] = r2;
  }
}

// Uniform-in-ball draw from explicit base randoms, then x = B y + c.
// Reference order (basic.py:376-381): y = (z / |z|) * u^(1/d).
__device__ __forceinline__ void ball_from_normals(double* z, int de, double u) {
  double n2 = 0.0;
  for (int j = 0; j < de; ++j) n2 = fma(z[j], z[j], n2);
  const double nrm = sqrt(n2);
  const double rad = pow(u, 1.0 / (double)de);
  for (int j = 0; j < de; ++j) z[j] = (z[j] / nrm) * rad;
}

__global__ void k_ell_sample_from(const double* __restrict__ zin,
                                  const double* __restrict__ u, int64_t n,
                                  int d, const double* __restrict__ c,
                                  const double* __restrict__ B,
                                  double* __restrict__ out) {
  extern __shared__ double sm[];
  const int stride = row_stride(d);
  double* rowA = sm;
  double* rowB = sm + (size_t)blockDim.x * stride;
  const int64_t base = (int64_t)blockIdx.x * blockDim.x;
  const int nrows = (int)min((int64_t)blockDim.x, n - base);
  load_rows(zin, base, nrows, d, stride, rowA);
  __syncthreads();
  if ((int)threadIdx.x < nrows) {
    double* z = rowA + threadIdx.x * stride;
    double* x = rowB + threadIdx.x * stride;
    ball_from_normals(z, d, u[base + threadIdx.x]);
    matvec_rows(B, d, 0, z, [&](int i, double v) { x[i] = v + __ldg(c + i); });
  }
  __syncthreads();
  store_rows(out, base, nrows, d, stride, rowB);
}

// ==========================================================================
// Union: count / contains (nautilus/bounds/union.py:269-289, 316-317)
// ==========================================================================

// mask semantics everywhere: a point is active iff mask == nullptr or
// mask[i] == mask_val (mask_val 4 lets a disposition array act as a mask).
__global__ void k_union_count(const int32_t* __restrict__ meta, int rec_off,
                              const double* __restrict__ data,
                              const double* __restrict__ points,
                              const uint8_t* __restrict__ mask, int mask_val,
                              int64_t n, int32_t* __restrict__ count,
                              uint8_t* __restrict__ contains,
                              uint8_t* __restrict__ passf) {
  extern __shared__ double sm[];
  const Rec rec{meta + rec_off};
  const int d = rec.d();
  const int stride = row_stride(d);
  double* rowA = sm;
  double* rowB = sm + (size_t)blockDim.x * stride;
  const int64_t base = (int64_t)blockIdx.x * blockDim.x;
  const int nrows = (int)min((int64_t)blockDim.x, n - base);
  load_rows(points, base, nrows, d, stride, rowA);
  __syncthreads();
  if ((int)threadIdx.x >= nrows) return;
  const int64_t i = base + threadIdx.x;
  const bool active = !mask || mask[i] == mask_val;
  int cnt = 0;
  bool in = false;
  if (active) {
    const double* x = rowA + threadIdx.x * stride;
    if (rec.kind() == 0) {
      in = cube_ok(x, nullptr, d);
      cnt = in ? 1 : 0;
    } else {
      cnt = union_count(rec, data, x, rowB + threadIdx.x * stride);
      in = cnt > 0;
      if (in && rec.unit()) in = cube_ok(x, nullptr, d);
    }
  }
  if (count) count[i] = cnt;
  if (contains) contains[i] = in ? 1 : 0;
  if (passf) passf[i] = (rec.kind() == 0 || rec.J() == 0) ? 1 : 0;
}

// ==========================================================================
// Proposal kernel: one Union.sample pass (nautilus/bounds/union.py:305-319),
// or UnitCube.sample (nautilus/bounds/basic.py:85) for a cube record.
// ==========================================================================

template <bool TEST>
__global__ void k_union_propose(
    const int32_t* __restrict__ meta, int rec_off,
    const double* __restrict__ data, int64_t n, uint64_t seed,
    uint64_t offset, uint32_t stream_id, const int32_t* __restrict__ k_in,
    const double* __restrict__ z_in, const double* __restrict__ cube_u_in,
    const double* __restrict__ u_in, const double* __restrict__ r_in,
    double* __restrict__ points, uint8_t* __restrict__ code,
    int32_t* __restrict__ n_bound_out,
    const unsigned long long* __restrict__ gather) {
  // gather != nullptr (nb200_materialize): proposal i is the GLOBAL proposal
  // index gather[i]; only the row is produced (same arithmetic as the cycle's)
  extern __shared__ double sm[];
  const Rec rec{meta + rec_off};
  const int d = rec.d();
  const int stride = row_stride(d);
  double* rowA = sm;                                  // x
  double* rowB = sm + (size_t)blockDim.x * stride;    // z / scratch
  const int64_t base = (int64_t)blockIdx.x * blockDim.x;
  const int nrows = (int)min((int64_t)blockDim.x, n - base);

  if ((int)threadIdx.x < nrows) {
    const int64_t i = base + threadIdx.x;
    double* x = rowA + threadIdx.x * stride;
    double* z = rowB + threadIdx.x * stride;
    const Philox rng(gather ? (uint64_t)gather[i] : offset + (uint64_t)i,
                     stream_id, seed);
    uint8_t cd = NB200_CODE_IN_SHELL;
    int nbnd = 0;

    if (rec.kind() == 0) {
      // UnitCube.sample: d uniforms in [0, 1)
      for (int j = 0; j < d; j += 2) {
        const uint4 w = rng.block(1 + (j >> 1));
        x[j] = u01_53(w.x, w.y);
        if (j + 1 < d) x[j + 1] = u01_53(w.z, w.w);
      }
      if (TEST) {
        for (int j = 0; j < d; ++j) x[j] = cube_u_in[i * d + j];
      }
    } else {
      int k;
      double r, u;
      if (TEST) {
        k = k_in[i];
        r = r_in[i];
        u = u_in[i];
      } else {
        const uint4 w = rng.block(0);
        const double uk = u01_32(w.x);
        const double* cdf = data + rec.off_cdf();
        const int K = rec.K();
        k = 0;
        while (k < K - 1 && !(uk < __ldg(cdf + k))) ++k;
        r = u01_32(w.y);
        u = u01_53(w.z, w.w);
      }
      const int32_t* m = rec.mix(k);
      const int de = m[0], nc = m[1];
      const int32_t* idx = rec.r + m[2];
      const uint32_t nblk = (uint32_t)((de + 3) >> 2);
      // cube dimensions first (basic.py:636), then the ellipsoid (:639)
      for (int q = 0; q < nc; q += 2) {
        double v0, v1 = 0.0;
        if (TEST) {
          v0 = cube_u_in[i * d + q];
          if (q + 1 < nc) v1 = cube_u_in[i * d + q + 1];
        } else {
          const uint4 w = rng.block(1 + nblk + (q >> 1));
          v0 = u01_53(w.x, w.y);
          v1 = u01_53(w.z, w.w);
        }
        x[idx[de + q]] = v0;
        if (q + 1 < nc) x[idx[de + q + 1]] = v1;
      }
      if (de > 0) {
        if (TEST) {
          for (int j = 0; j < de; ++j) z[j] = z_in[i * d + j];
        } else {
          for (int j = 0; j < de; j += 4) {
            const uint4 w = rng.block(1 + (j >> 2));
            float n0, n1, n2, n3;
            normal2(w.x, w.y, n0, n1);
            normal2(w.z, w.w, n2, n3);
            z[j] = (double)n0;
            if (j + 1 < de) z[j + 1] = (double)n1;
            if (j + 2 < de) z[j + 2] = (double)n2;
            if (j + 3 < de) z[j + 3] = (double)n3;
          }
        }
        ball_from_normals(z, de, u);
        const double* c = data + m[3];
        // B is exactly lower-triangular (Cholesky factor, basic.py:308)
        matvec_rows(data + m[4], de, 1, z, [&](int ii, double v) {
          x[idx[ii]] = v + __ldg(c + ii);
        });
      }
      // unit-cube filter (union.py:313-314), overlap count (:316-317),
      // accept iff r > 1 - 1/n_bound (:318-319; n_bound == 0 -> -inf -> accept)
      if (gather) {
        // row only
      } else if (rec.unit() && !cube_ok(x, nullptr, d)) {
        cd = NB200_CODE_CUBE_REJECT;
      } else {
        nbnd = union_count(rec, data, x, z);
        const double p = 1.0 - 1.0 / (double)nbnd;
        if (!(r > p)) cd = NB200_CODE_OVERLAP_REJECT;
      }
    }
    if (code) code[i] = cd;
    if (n_bound_out) n_bound_out[i] = nbnd;
  }
  __syncthreads();
  store_rows(points, base, nrows, d, stride, rowA);
}

// ==========================================================================
// NeuralBound (nautilus/bounds/neural.py:115-126)
// ==========================================================================

// For active candidates: in_ell = Ellipsoid.contains(x); t = transform(x).
// Writes t rows (whitened coordinates) and maskj = active & in_ell.  If the
// neural bound has no emulator (n_networks == 0) membership is in_ell itself.
__global__ void k_neural_prep(const int32_t* __restrict__ meta, int rec_off,
                              const double* __restrict__ data, int j,
                              const double* __restrict__ points,
                              const uint8_t* __restrict__ cand, int cand_val,
                              int64_t n, double* __restrict__ t_rows,
                              uint8_t* __restrict__ maskj,
                              uint8_t* __restrict__ passf) {
  extern __shared__ double sm[];
  const Rec rec{meta + rec_off};
  const int d = rec.d();
  const int stride = row_stride(d);
  double* rowA = sm;
  double* rowB = sm + (size_t)blockDim.x * stride;
  const int64_t base = (int64_t)blockIdx.x * blockDim.x;
  const int nrows = (int)min((int64_t)blockDim.x, n - base);
  load_rows(points, base, nrows, d, stride, rowA);
  __syncthreads();
  const int32_t* nb = rec.nb(j);
  if ((int)threadIdx.x < nrows) {
    const int64_t i = base + threadIdx.x;
    const bool active = !cand || cand[i] == cand_val;
    bool in = false;
    double* x = rowA + threadIdx.x * stride;
    double* s = rowB + threadIdx.x * stride;
    if (active) {
      // t overwrites x element i only after s has been formed from all of x
      const double r2 = whiten_r2(x, nullptr, d, data + nb[0], data + nb[1],
                                  nb[2], s, x);
      in = r2 < 1.0;
    }
    maskj[i] = in ? 1 : 0;
    if (nb[3] == 0 && in && passf) passf[i] = 1;
  }
  __syncthreads();
  if (t_rows) store_rows(t_rows, base, nrows, d, stride, rowA);
}

// fp64 emulator: NeuralNetworkEmulator.predict (nautilus/neural.py:114-116)
// = mean over networks of the sklearn MLP forward pass on (t - mean) / scale.
// Rows: X (d), H0, H1 (max hidden width).
__global__ void k_mlp_f64(const int32_t* __restrict__ meta, int rec_off,
                          const double* __restrict__ data, int j,
                          const double* __restrict__ t_rows,
                          const uint8_t* __restrict__ mask, int64_t n,
                          int wstride, double* __restrict__ score_out,
                          uint8_t* __restrict__ passf) {
  extern __shared__ double sm[];
  const Rec rec{meta + rec_off};
  const int d = rec.d();
  const int stride = row_stride(d);
  double* rowX = sm;
  double* rowH0 = rowX + (size_t)blockDim.x * stride;
  double* rowH1 = rowH0 + (size_t)blockDim.x * wstride;
  const int64_t base = (int64_t)blockIdx.x * blockDim.x;
  const int nrows = (int)min((int64_t)blockDim.x, n - base);
  load_rows(t_rows, base, nrows, d, stride, rowX);
  __syncthreads();
  if ((int)threadIdx.x >= nrows) return;
  const int64_t i = base + threadIdx.x;
  if (mask && !mask[i]) {
    if (score_out) score_out[i] = nan("");
    return;
  }
  const int32_t* nb = rec.nb(j);
  const int n_net = nb[3], n_lay = nb[4];
  const double* mean = data + nb[5];
  const double* scale = data + nb[6];
  const int32_t* sizes = rec.r + nb[8];
  const int32_t* wtab = rec.r + nb[9];
  double* x = rowX + threadIdx.x * stride;
  for (int k = 0; k < d; ++k) x[k] = (x[k] - __ldg(mean + k)) / __ldg(scale + k);

  double sum = 0.0;
  for (int net = 0; net < n_net; ++net) {
    const double* a = x;
    double* h = rowH0 + threadIdx.x * wstride;
    double* h_other = rowH1 + threadIdx.x * wstride;
    double y = 0.0;
    for (int l = 0; l < n_lay; ++l) {
      const int fi = sizes[l], fo = sizes[l + 1];
      const double* W = data + wtab[(net * n_lay + l) * 2];
      const double* b = data + wtab[(net * n_lay + l) * 2 + 1];
      const bool last = (l == n_lay - 1);
      for (int n0 = 0; n0 < fo; n0 += 8) {
        double acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.0;
        const int nq = min(8, fo - n0);
        if (nq == 8) {
          for (int k = 0; k < fi; ++k) {
            const double ak = a[k];
            const double* w = W + (size_t)k * fo + n0;
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = fma(__ldg(w + q), ak, acc[q]);
          }
        } else {
          for (int k = 0; k < fi; ++k) {
            const double ak = a[k];
            const double* w = W + (size_t)k * fo + n0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q < nq) acc[q] = fma(__ldg(w + q), ak, acc[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q < nq) {
            double v = acc[q] + __ldg(b + n0 + q);
            if (!last) v = fmax(v, 0.0);
            if (last) y = v; else h[n0 + q] = v;
          }
        }
      }
      a = h;
      double* tmp = h; h = h_other; h_other = tmp;
    }
    sum = (net == 0) ? y : sum + y;
  }
  const double score = sum / (double)n_net;
  if (score_out) score_out[i] = score;
  if (passf) {
    const double thr = __ldg(data + nb[7]);   // score_predict_min - 1e-9
    if (score > thr) passf[i] = 1;
  }
}

// ==========================================================================
// element-wise glue
// ==========================================================================

// mode 0: cand = (code == 4), passf = 0          (start of the NN filter)
// mode 1: code 4 & !passf          -> NN_REJECT  (nautilus.py:217-219)
// mode 2: code 4 & cand & passf    -> EXCLUDED   (sampler.py:796-801)
// mode 3: out = cand & passf                     (NautilusBound.contains)
// mode 4: out = passf restricted to active mask  (neural filter only)
__global__ void k_apply(int mode, int64_t n, uint8_t* __restrict__ code,
                        uint8_t* __restrict__ cand,
                        uint8_t* __restrict__ passf,
                        uint8_t* __restrict__ out,
                        const uint8_t* __restrict__ mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  switch (mode) {
    case 0:
      cand[i] = code[i] == NB200_CODE_IN_SHELL ? 1 : 0;
      passf[i] = 0;
      break;
    case 1:
      if (code[i] == NB200_CODE_IN_SHELL && !passf[i])
        code[i] = NB200_CODE_NN_REJECT;
      break;
    case 2:
      if (code[i] == NB200_CODE_IN_SHELL && cand[i] && passf[i])
        code[i] = NB200_CODE_EXCLUDED;
      break;
    case 3:
      out[i] = (cand[i] && passf[i]) ? 1 : 0;
      break;
    case 4:
      out[i] = ((!mask || mask[i]) && passf[i]) ? 1 : 0;
      break;
  }
}

__global__ void k_fill_u8(uint8_t* p, int64_t n, uint8_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ==========================================================================
// Built-in likelihoods
// ==========================================================================

__global__ void k_loglike(const double* __restrict__ points,
                          const uint8_t* __restrict__ code, int64_t n, int d,
                          int like_id, const double* __restrict__ p,
                          double* __restrict__ log_l) {
  extern __shared__ double sm[];
  const int stride = row_stride(d);
  const int64_t base = (int64_t)blockIdx.x * blockDim.x;
  const int nrows = (int)min((int64_t)blockDim.x, n - base);
  load_rows(points, base, nrows, d, stride, sm);
  __syncthreads();
  if ((int)threadIdx.x >= nrows) return;
  const int64_t i = base + threadIdx.x;
  if (code && code[i] != NB200_CODE_IN_SHELL) {
    log_l[i] = nan("");
    return;
  }
  log_l[i] = loglike_eval(like_id, p, sm + threadIdx.x * stride, d);
}

// ==========================================================================
// Shell reductions (nautilus/sampler.py:925-943, 1144)
// ==========================================================================

__global__ void __launch_bounds__(STAT_THREADS)
k_stats_partial(const double* __restrict__ log_l,
                const uint8_t* __restrict__ code, int64_t n, double log_l_min,
                StatPartial* __restrict__ partial) {
  Lse acc;
  acc.init();
  long long cnt[NB200_N_CNT];
#pragma unroll
  for (int q = 0; q < NB200_N_CNT; ++q) cnt[q] = 0;
  // contiguous chunk per block, strided by thread inside: fixed shape
  const int64_t per_block = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per_block;
  const int64_t hi = min(n, lo + per_block);
  for (int64_t i = lo + threadIdx.x; i < hi; i += STAT_THREADS) {
    const int cd = code ? code[i] : NB200_CODE_IN_SHELL;
    cnt[NB200_CNT_RAW] += 1;
    if (cd == NB200_CODE_IN_SHELL) {
      const double l = log_l ? log_l[i] : 0.0;
      acc.add(l);
      cnt[NB200_CNT_IN_SHELL] += 1;
      if (l >= log_l_min) cnt[NB200_CNT_UPDATE] += 1;
    } else {
      cnt[1 + cd] += 1;  // codes 0..3 -> counters 1..4
    }
  }
  stat_block_reduce<STAT_THREADS>(acc, cnt, partial + blockIdx.x);
}

__global__ void __launch_bounds__(STAT_THREADS)
k_stats_final(const StatPartial* __restrict__ partial, int nblocks,
              double* __restrict__ lse, long long* __restrict__ counters) {
  Lse acc;
  acc.init();
  long long cnt[NB200_N_CNT];
#pragma unroll
  for (int q = 0; q < NB200_N_CNT; ++q) cnt[q] = 0;
  for (int b = threadIdx.x; b < nblocks; b += STAT_THREADS) {
    Lse o;
    o.m = partial[b].m; o.s1 = partial[b].s1; o.s2 = partial[b].s2;
    acc.merge(o);
#pragma unroll
    for (int q = 0; q < NB200_N_CNT; ++q) cnt[q] += partial[b].cnt[q];
  }
  __shared__ StatPartial tot;
  stat_block_reduce<STAT_THREADS>(acc, cnt, &tot);
  __syncthreads();
  if (threadIdx.x == 0) {
    lse[0] = tot.m; lse[1] = tot.s1; lse[2] = tot.s2; lse[3] = 0.0;
    for (int q = 0; q < NB200_N_CNT; ++q) counters[q] = tot.cnt[q];
  }
}

// ==========================================================================
// Stable compaction of in-shell rows
// ==========================================================================

constexpr int CMP_THREADS = 256;
constexpr int CMP_ITEMS = 1024;  // items per block

__global__ void __launch_bounds__(CMP_THREADS)
k_compact_count(const uint8_t* __restrict__ code, int64_t n,
                long long* __restrict__ block_count) {
  const int64_t lo = (int64_t)blockIdx.x * CMP_ITEMS;
  int c = 0;
  for (int p = 0; p < CMP_ITEMS / CMP_THREADS; ++p) {
    const int64_t i = lo + p * CMP_THREADS + threadIdx.x;
    if (i < n && code[i] == NB200_CODE_IN_SHELL) ++c;
  }
  c = warp_sum(c);
  __shared__ int sm[CMP_THREADS / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < CMP_THREADS / 32; ++w) t += sm[w];
    block_count[blockIdx.x] = t;
  }
}

// exclusive scan of block counts by one block (nblocks <= ~16k for 2^24)
__global__ void __launch_bounds__(1024)
k_compact_scan(long long* __restrict__ block_count, int nblocks,
               long long* __restrict__ total) {
  __shared__ long long sm[1024];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int start = 0; start < nblocks; start += 1024) {
    const int b = start + threadIdx.x;
    const long long v = b < nblocks ? block_count[b] : 0;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const long long add = threadIdx.x >= o ? sm[threadIdx.x - o] : 0;
      __syncthreads();
      sm[threadIdx.x] += add;
      __syncthreads();
    }
    if (b < nblocks) block_count[b] = carry + sm[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sm[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(CMP_THREADS)
k_compact_scatter(const double* __restrict__ points,
                  const double* __restrict__ log_l,
                  const uint8_t* __restrict__ code, int64_t n, int d,
                  const long long* __restrict__ block_off,
                  double* __restrict__ out_points,
                  double* __restrict__ out_log_l) {
  __shared__ int warp_cnt[CMP_THREADS / 32];
  __shared__ long long running;
  __shared__ long long dest[CMP_THREADS];
  const int64_t lo = (int64_t)blockIdx.x * CMP_ITEMS;
  if (threadIdx.x == 0) running = block_off[blockIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int p = 0; p < CMP_ITEMS / CMP_THREADS; ++p) {
    const int64_t i = lo + p * CMP_THREADS + threadIdx.x;
    const bool keep = i < n && code[i] == NB200_CODE_IN_SHELL;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) before += warp_cnt[w];
    dest[threadIdx.x] = keep ? running + before : -1;
    if (keep && out_log_l) out_log_l[running + before] = log_l[i];
    __syncthreads();
    // cooperative, coalesced row copy: one warp per kept row
    if (out_points) {
      for (int r = warp; r < CMP_THREADS; r += CMP_THREADS / 32) {
        const long long dst = dest[r];
        if (dst < 0) continue;
        const double* src = points + (lo + p * CMP_THREADS + r) * (int64_t)d;
        double* o = out_points + dst * (int64_t)d;
        for (int j = lane; j < d; j += 32) o[j] = src[j];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < CMP_THREADS / 32; ++w) t += warp_cnt[w];
      running += t;
    }
    __syncthreads();
  }
}

// Index form of the compaction: for every in-shell proposal its GLOBAL proposal
// index (offset + i, the Philox counter that regenerates the row, see
// nb200_materialize) and its log_l -- 16 bytes per kept proposal instead of
// the 8 d + 8 of a row.  Every block derives its own exclusive prefix from the
// block counts (<= n / 1024 values, L2-resident), so there is no scan launch;
// the last block also writes the total.
__global__ void __launch_bounds__(CMP_THREADS)
k_compact_index(const double* __restrict__ log_l,
                const uint8_t* __restrict__ code, int64_t n,
                unsigned long long offset,
                const long long* __restrict__ block_count,
                unsigned long long* __restrict__ out_index,
                double* __restrict__ out_log_l, long long* __restrict__ total) {
  __shared__ int warp_cnt[CMP_THREADS / 32];
  __shared__ long long warp_sum_sm[CMP_THREADS / 32];
  __shared__ long long running;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long pre = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += CMP_THREADS)
    pre += block_count[b];
  pre = warp_sum(pre);
  if (lane == 0) warp_sum_sm[warp] = pre;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < CMP_THREADS / 32; ++w) t += warp_sum_sm[w];
    running = t;
    if (blockIdx.x == gridDim.x - 1) *total = t + block_count[blockIdx.x];
  }
  __syncthreads();
  const int64_t lo = (int64_t)blockIdx.x * CMP_ITEMS;
  for (int p = 0; p < CMP_ITEMS / CMP_THREADS; ++p) {
    const int64_t i = lo + p * CMP_THREADS + threadIdx.x;
    const bool keep = i < n && code[i] == NB200_CODE_IN_SHELL;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = __popc(bal & ((1u << lane) - 1u));
    int all = 0;
    for (int w = 0; w < CMP_THREADS / 32; ++w) {
      if (w < warp) before += warp_cnt[w];
      all += warp_cnt[w];
    }
    if (keep) {
      const long long dst = running + before;
      out_index[dst] = offset + (unsigned long long)i;
      if (out_log_l) out_log_l[dst] = log_l[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) running += all;
    __syncthreads();
  }
}

// ==========================================================================
// host-side launch helpers
// ==========================================================================

struct Workspace {        // carve-up of the caller-provided scratch
  double* t_rows;         // n * d
  uint8_t* cand;          // n
  uint8_t* passf;         // n
  uint8_t* maskj;         // n
  float* xs32;            // n * round8(d): standardised tf32 rows
  StatPartial* partial;   // STAT_MAX_BLOCKS
  long long* block_count; // n / CMP_ITEMS + 2
  long long* total;       // 1
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t workspace_layout(int64_t n, int d, char* base, Workspace* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  Workspace w;
  w.t_rows = (double*)take((size_t)n * d * sizeof(double));
  w.cand = (uint8_t*)take((size_t)n);
  w.passf = (uint8_t*)take((size_t)n);
  w.maskj = (uint8_t*)take((size_t)n);
  w.xs32 = (float*)take((size_t)n * ((d + 8) / 8 * 8) * sizeof(float));
  w.partial = (StatPartial*)take(sizeof(StatPartial) * STAT_MAX_BLOCKS);
  w.block_count =
      (long long*)take(sizeof(long long) * ((size_t)n / CMP_ITEMS + 2));
  w.total = (long long*)take(sizeof(long long));
  if (ws) *ws = w;
  return off;
}

static int launch_apply(int mode, int64_t n, uint8_t* code, uint8_t* cand,
                        uint8_t* passf, uint8_t* out, const uint8_t* mask,
                        cudaStream_t st) {
  if (n == 0) return 0;
  ProfScope prof(ST_GLUE, st);
  k_apply<<<blocks_for(n, 256), 256, 0, st>>>(mode, n, code, cand, passf, out,
                                              mask);
  NB_LAUNCH_OK();
  return 0;
}

int launch_mlp_tf32(const int32_t* meta_h, const double* data_d, int bound,
                    int j, const double* t_rows, const uint8_t* mask,
                    int64_t n, double* score_out, uint8_t* passf,
                    float* xs32_ws, int mode, cudaStream_t st);

int launch_mlp_tf32_rows(const int32_t* meta_h, const double* data_d,
                         int bound, int j, const float* xs32,
                         const uint8_t* mask, int64_t n, uint8_t* code,
                         double log_l_min, double* log_l, void* partial,
                         int* n_partial_out, int mode, cudaStream_t st);
bool mlp_tf32_resident(const int32_t* meta_h, int bound, int j, int mode);
const int32_t* tc_header_words(const int32_t* meta_h, int bound, int j,
                               int mode);
bool front_applicable(const int32_t* meta_h, int bound, size_t* smem_out,
                      struct FrontArgs* args);
int launch_front(const int32_t* meta_h, const int32_t* meta_d,
                 const double* data_d, int bound, int64_t n, uint64_t seed,
                 uint64_t offset, uint32_t stream_id, double* points,
                 uint8_t* code, uint8_t* maskj, float* xs32, int like_id,
                 const double* like_p, double* log_l,
                 const unsigned long long* gather, int xs_f16,
                 cudaStream_t st);

// emulator of neural bound j on whitened rows; ORs into passf and/or writes
// the scores.
static int launch_mlp(const int32_t* meta_h, const int32_t* meta_d,
                      const double* data_d, int bound, int j,
                      const double* t_rows, const uint8_t* mask, int64_t n,
                      double* score_out, uint8_t* passf, int mlp_mode,
                      float* xs32_ws, cudaStream_t st) {
  if (n == 0) return 0;
  const Rec rec = record(meta_h, bound);
  const int32_t* nb = rec.nb(j);
  NB_CHECK(nb[3] > 0, "neural bound has no emulator");
  ProfScope prof(ST_MLP, st);
  if (mode_tc(mlp_mode))
    return launch_mlp_tf32(meta_h, data_d, bound, j, t_rows, mask, n,
                           score_out, passf, xs32_ws, mlp_mode, st);
  NB_CHECK(mlp_mode == NB200_MLP_F64, "unknown mlp_mode");
  const int d = rec.d();
  const int32_t* sizes = rec.r + nb[8];
  int wmax = 1;
  for (int l = 1; l < nb[4]; ++l) wmax = sizes[l] > wmax ? sizes[l] : wmax;
  const int wstride = wmax | 1;
  const int threads = wmax <= 128 ? 64 : 32;
  const size_t smem =
      (size_t)threads * ((d | 1) + 2 * wstride) * sizeof(double);
  NB_CHECK(smem <= 227 * 1024, "emulator too wide for shared memory");
  if (opt_in_smem(k_mlp_f64, smem)) return 2;
  k_mlp_f64<<<blocks_for(n, threads), threads, smem, st>>>(
      meta_d, (int)(rec.r - meta_h), data_d, j, t_rows, mask, n, wstride,
      score_out, passf);
  NB_LAUNCH_OK();
  return 0;
}

// any_j NeuralBound_j.contains over candidates; passf must be pre-initialised
// (0, or 1 where J == 0).  cand may be a code array (cand_val 4) or a mask.
static int neural_any(const int32_t* meta_h, const int32_t* meta_d,
                      const double* data_d, int bound, const double* points,
                      const uint8_t* cand, int cand_val, int64_t n,
                      const Workspace& ws, int mlp_mode, cudaStream_t st) {
  const Rec rec = record(meta_h, bound);
  if (rec.kind() == 0 || n == 0) return 0;
  const int d = rec.d();
  const int threads = threads_for(d);
  const size_t smem = smem_rows(threads, d, 2);
  if (opt_in_smem(k_neural_prep, smem)) return 2;
  for (int j = 0; j < rec.J(); ++j) {
    const bool has_emu = rec.nb(j)[3] > 0;
    {
    ProfScope prof(ST_PREP, st);
    k_neural_prep<<<blocks_for(n, threads), threads, smem, st>>>(
        meta_d, (int)(rec.r - meta_h), data_d, j, points, cand, cand_val, n,
        has_emu ? ws.t_rows : nullptr, ws.maskj, ws.passf);
    NB_LAUNCH_OK();
    }
    if (has_emu) {
      const int rc = launch_mlp(meta_h, meta_d, data_d, bound, j, ws.t_rows,
                                ws.maskj, n, nullptr, ws.passf, mlp_mode,
                                ws.xs32, st);
      if (rc) return rc;
    }
  }
  return 0;
}

static int union_contains_launch(const int32_t* meta_h, const int32_t* meta_d,
                                 const double* data_d, int bound,
                                 const double* points, const uint8_t* mask,
                                 int mask_val, int64_t n, int32_t* count,
                                 uint8_t* contains, uint8_t* passf,
                                 cudaStream_t st) {
  if (n == 0) return 0;
  const Rec rec = record(meta_h, bound);
  const int d = rec.d();
  const int threads = threads_for(d);
  const size_t smem = smem_rows(threads, d, 2);
  if (opt_in_smem(k_union_count, smem)) return 2;
  ProfScope prof(ST_UNION, st);
  k_union_count<<<blocks_for(n, threads), threads, smem, st>>>(
      meta_d, (int)(rec.r - meta_h), data_d, points, mask, mask_val, n, count,
      contains, passf);
  NB_LAUNCH_OK();
  return 0;
}

// ---- grouped later-bound exclusion (csrc/nb200_exclude.cu) ------------------
int launch_excl_pairs(const int32_t* meta_d, int first_later, int n_later,
                      int f16, PairRec* pairs, int* pair_base,
                      cudaStream_t st);
size_t excl_prep_smem(int d, int k0p);
int launch_excl_prep(const int32_t* meta_d, const double* data_d,
                     int first_later, int n_later, int d, int k0p, int f16,
                     const double* points, const unsigned long long* cand_idx,
                     const unsigned long long* n_cand, long long chunk_lo,
                     long long chunk_cap, const int* pair_base,
                     unsigned int* seg_count, float* xs, unsigned int* cid,
                     long long seg_stride, uint8_t* excl, cudaStream_t st);
int launch_excl_apply(const unsigned long long* cand_idx,
                      const unsigned long long* n_cand, const uint8_t* excl,
                      uint8_t* code, int64_t n, cudaStream_t st);
int run_mlp_tf32_grouped(const int32_t* hdr32, const float* xs_segments,
                         const TcGroupArgs& G, cudaStream_t st);

constexpr int64_t EXCL_CHUNK = 1 << 18;   // candidates per pass (recommended)

static inline int k0p_of(int d) { return (d + 1 + 7) / 8 * 8; }

struct ExclWs {
  unsigned long long* cand_idx;   // n
  uint8_t* excl;                  // n
  PairRec* pairs;                 // P
  int* pair_base;                 // n_later
  unsigned int* seg_count;        // P
  float* xs;                      // P * cap * k0p
  unsigned int* cid;              // P * cap
  int64_t cap;                    // rows per segment (candidates per pass)
};

// carve the exclusion scratch out of [base, base + avail); cap < 0 asks for
// the size that holds `want_cap` candidates per pass
static size_t excl_layout(int64_t n, int P, int n_later, int k0p, char* base,
                          size_t avail, int64_t want_cap, ExclWs* out) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  ExclWs w;
  w.cand_idx = (unsigned long long*)take(sizeof(unsigned long long) * n);
  w.excl = (uint8_t*)take((size_t)n);
  w.pairs = (PairRec*)take(sizeof(PairRec) * (size_t)(P > 0 ? P : 1));
  w.pair_base = (int*)take(sizeof(int) * (size_t)n_later);
  w.seg_count = (unsigned int*)take(sizeof(unsigned int) * (size_t)(P > 0 ? P : 1));
  const size_t row = (size_t)(P > 0 ? P : 1) * ((size_t)k0p * 4 + 4);
  int64_t cap = want_cap;
  if (cap < 0) {
    cap = off + 512 < avail ? (int64_t)((avail - off - 512) / row) : 0;
    cap = cap / 128 * 128;
    const int64_t n128 = (n + 127) / 128 * 128;
    if (cap > n128) cap = n128;
  }
  w.cap = cap;
  w.xs = (float*)take((size_t)(P > 0 ? P : 1) * (size_t)cap * k0p * 4);
  w.cid = (unsigned int*)take((size_t)(P > 0 ? P : 1) * (size_t)cap * 4);
  if (out) *out = w;
  return off;
}

// can records first_later .. +n_later be handled by the grouped pass?
// *n_pairs = (later bound, neural bound) pairs; *hdr = the tensor-core header
// shared by every pair that has an emulator (nullptr: none has one)
static bool excl_applicable(const int32_t* meta_h, int first_later,
                            int n_later, int d, int mode, int* n_pairs,
                            const int32_t** hdr) {
  const char* e = getenv("NB200_EXCLUDE");
  if (e && strcmp(e, "loop") == 0) return false;
  if (d > 48) return false;
  int P = 0;
  const int32_t* first = nullptr;
  for (int l = first_later; l < first_later + n_later; ++l) {
    const Rec rec = record(meta_h, l);
    if (rec.d() != d) return false;
    if (rec.kind() == 0) continue;
    for (int j = 0; j < rec.J(); ++j, ++P) {
      const int32_t* nb = rec.nb(j);
      if (nb[3] <= 0) continue;                     // ellipsoid only
      const int32_t* h = tc_header_words(meta_h, l, j, mode);
      if (!h) return false;                // no resident tensor-core blob
      // one architecture for all pairs (words 20..23 of the fp16 header
      // hold the blob's position, words 30, 31 the threshold)
      if (!first) first = h;
      else if (memcmp(first, h, 20 * sizeof(int32_t)) != 0 ||
               memcmp(first + 24, h + 24, 6 * sizeof(int32_t)) != 0)
        return false;
    }
  }
  if (P > 1 << 15) return false;
  *n_pairs = P;
  *hdr = first;
  return true;
}

static int check_bound(const int32_t* meta_h, int bound) {
  NB_CHECK(meta_h != nullptr, "meta_h is NULL");
  NB_CHECK(bound >= 0 && bound < meta_h[0], "bound index out of range");
  const Rec rec = record(meta_h, bound);
  NB_CHECK(rec.d() >= 1 && rec.d() <= NB200_D_MAX, "n_dim out of range");
  return 0;
}

}  // namespace nb200

using namespace nb200;

// ==========================================================================
// C ABI
// ==========================================================================

extern "C" {

const char* nb200_last_error(void) { return g_err; }
int nb200_version(void) { return NB200_VERSION; }
int64_t nb200_launch_count(void) { return g_launches; }

void nb200_profile_enable(int on) {
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on = on != 0;
}

int nb200_profile_collect(double* ms_by_stage, int64_t* calls_by_stage) {
  NB_CUDA(cudaDeviceSynchronize());
  for (int s = 0; s < N_STAGES; ++s) { ms_by_stage[s] = 0; calls_by_stage[s] = 0; }
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ms_by_stage[r.stage] += ms;
      calls_by_stage[r.stage] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return 0;
}

const char* nb200_profile_stage_name(int stage) {
  static const char* names[N_STAGES] = {
      "union_propose", "union_count", "neural_prep", "mlp_predict", "glue",
      "loglike", "stats", "compact", "fused_cycle", "mlp_fit"};
  return (stage >= 0 && stage < N_STAGES) ? names[stage] : "";
}

int nb200_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  NB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  NB_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

size_t nb200_workspace_bytes(int64_t n, int d) {
  return workspace_layout(n < 1 ? 1 : n, d < 1 ? 1 : d, nullptr, nullptr);
}

size_t nb200_cycle_workspace_bytes(int64_t n, int d, int n_pairs) {
  if (n < 1) n = 1;
  if (d < 1) d = 1;
  size_t bytes = nb200_workspace_bytes(n, d);
  if (n_pairs > 0) {
    int64_t chunk = EXCL_CHUNK;
    const char* e = getenv("NB200_EXCL_CHUNK");     // tests: force more passes
    if (e && atoll(e) >= 128) chunk = atoll(e);
    const int64_t cap = (n < chunk ? n : chunk) + 127;
    bytes += excl_layout(n, n_pairs, n_pairs, k0p_of(d), nullptr, 0,
                         cap / 128 * 128, nullptr) + 1024;
  }
  return bytes;
}

int nb200_ell_transform(const double* points_d, int64_t n, int d,
                        const double* c_d, const double* M_d, int inverse,
                        double* out_d, void* stream) {
  NB_CHECK(d >= 1 && d <= NB200_D_MAX, "n_dim out of range");
  NB_CHECK(n >= 0, "negative n");
  if (n == 0) return 0;
  const int threads = threads_for(d);
  const size_t smem = smem_rows(threads, d, 2);
  if (opt_in_smem(k_ell_transform, smem)) return 2;
  k_ell_transform<<<blocks_for(n, threads), threads, smem,
                    (cudaStream_t)stream>>>(points_d, n, d, c_d, M_d, inverse,
                                            out_d);
  NB_LAUNCH_OK();
  return 0;
}

int nb200_ell_contains(const double* points_d, int64_t n, int d,
                       const double* c_d, const double* Binv_d, uint8_t* out_d,
                       double* r2_d, void* stream) {
  NB_CHECK(d >= 1 && d <= NB200_D_MAX, "n_dim out of range");
  NB_CHECK(n >= 0, "negative n");
  if (n == 0) return 0;
  const int threads = threads_for(d);
  const size_t smem = smem_rows(threads, d, 2);
  if (opt_in_smem(k_ell_contains, smem)) return 2;
  k_ell_contains<<<blocks_for(n, threads), threads, smem,
                   (cudaStream_t)stream>>>(points_d, n, d, c_d, Binv_d, out_d,
                                           r2_d);
  NB_LAUNCH_OK();
  return 0;
}

int nb200_ell_sample_from(const double* z_d, const double* u_d, int64_t n,
                          int d, const double* c_d, const double* B_d,
                          double* out_d, void* stream) {
  NB_CHECK(d >= 1 && d <= NB200_D_MAX, "n_dim out of range");
  NB_CHECK(n >= 0, "negative n");
  if (n == 0) return 0;
  const int threads = threads_for(d);
  const size_t smem = smem_rows(threads, d, 2);
  if (opt_in_smem(k_ell_sample_from, smem)) return 2;
  k_ell_sample_from<<<blocks_for(n, threads), threads, smem,
                      (cudaStream_t)stream>>>(z_d, u_d, n, d, c_d, B_d, out_d);
  NB_LAUNCH_OK();
  return 0;
}

int nb200_union_count(const int32_t* meta_h, const int32_t* meta_d,
                      const double* data_d, int bound, const double* points_d,
                      const uint8_t* in_mask_d, int64_t n, int32_t* count_d,
                      uint8_t* contains_d, void* stream) {
  if (check_bound(meta_h, bound)) return 1;
  NB_CHECK(n >= 0, "negative n");
  return union_contains_launch(meta_h, meta_d, data_d, bound, points_d,
                               in_mask_d, 1, n, count_d, contains_d, nullptr,
                               (cudaStream_t)stream);
}

int nb200_union_propose(const int32_t* meta_h, const int32_t* meta_d,
                        const double* data_d, int bound, int64_t n,
                        uint64_t seed, uint64_t offset, uint32_t stream_id,
                        const int32_t* k_d, const double* z_d,
                        const double* cube_u_d, const double* u_d,
                        const double* r_d, double* points_d, uint8_t* code_d,
                        int32_t* n_bound_d, void* stream) {
  if (check_bound(meta_h, bound)) return 1;
  NB_CHECK(n >= 0, "negative n");
  if (n == 0) return 0;
  const Rec rec = record(meta_h, bound);
  const int d = rec.d();
  const int threads = threads_for(d);
  const size_t smem = smem_rows(threads, d, 2);
  const int rec_off = (int)(rec.r - meta_h);
  const bool test = (rec.kind() == 0) ? (cube_u_d != nullptr) : (k_d != nullptr);
  ProfScope prof(ST_PROPOSE, (cudaStream_t)stream);
  if (test) {
    if (rec.kind() != 0)
      NB_CHECK(z_d && cube_u_d && u_d && r_d,
               "test mode needs k, z, cube_u, u and r");
    if (opt_in_smem(k_union_propose<true>, smem)) return 2;
    k_union_propose<true><<<blocks_for(n, threads), threads, smem,
                            (cudaStream_t)stream>>>(
        meta_d, rec_off, data_d, n, seed, offset, stream_id, k_d, z_d,
        cube_u_d, u_d, r_d, points_d, code_d, n_bound_d, nullptr);
  } else {
    if (opt_in_smem(k_union_propose<false>, smem)) return 2;
    k_union_propose<false><<<blocks_for(n, threads), threads, smem,
                             (cudaStream_t)stream>>>(
        meta_d, rec_off, data_d, n, seed, offset, stream_id, nullptr, nullptr,
        nullptr, nullptr, nullptr, points_d, code_d, n_bound_d, nullptr);
  }
  NB_LAUNCH_OK();
  return 0;
}

int nb200_mlp_predict(const int32_t* meta_h, const int32_t* meta_d,
                      const double* data_d, int bound, int j,
                      const double* x_d, int64_t n, double* out_d,
                      int mlp_mode, void* workspace_d, size_t workspace_bytes,
                      void* stream) {
  if (check_bound(meta_h, bound)) return 1;
  const Rec rec = record(meta_h, bound);
  NB_CHECK(rec.kind() == 1 && j >= 0 && j < rec.J(), "neural bound index");
  NB_CHECK(n >= 0, "negative n");
  NB_CHECK(workspace_bytes >= nb200_workspace_bytes(n, rec.d()),
           "workspace too small");
  Workspace ws;
  workspace_layout(n < 1 ? 1 : n, rec.d(), (char*)workspace_d, &ws);
  return launch_mlp(meta_h, meta_d, data_d, bound, j, x_d, nullptr, n, out_d,
                    nullptr, mlp_mode, ws.xs32, (cudaStream_t)stream);
}

int nb200_bound_contains(const int32_t* meta_h, const int32_t* meta_d,
                         const double* data_d, int bound, int which,
                         const double* points_d, const uint8_t* in_mask_d,
                         int64_t n, uint8_t* out_d, int mlp_mode,
                         void* workspace_d, size_t workspace_bytes,
                         void* stream) {
  if (check_bound(meta_h, bound)) return 1;
  NB_CHECK(n >= 0, "negative n");
  NB_CHECK(which >= 0 && which <= 2, "which must be 0, 1 or 2");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const Rec rec = record(meta_h, bound);
  NB_CHECK(workspace_bytes >= nb200_workspace_bytes(n, rec.d()),
           "workspace too small");
  Workspace ws;
  workspace_layout(n, rec.d(), (char*)workspace_d, &ws);
  if (which == 2) {
    NB_CHECK(rec.kind() == 1, "neural filter needs a nautilus bound");
    k_fill_u8<<<blocks_for(n, 256), 256, 0, st>>>(ws.passf, n,
                                                  rec.J() == 0 ? 1 : 0);
    NB_LAUNCH_OK();
    int rc = neural_any(meta_h, meta_d, data_d, bound, points_d, in_mask_d, 1,
                        n, ws, mlp_mode, st);
    if (rc) return rc;
    return launch_apply(4, n, nullptr, nullptr, ws.passf, out_d, in_mask_d, st);
  }
  int rc = union_contains_launch(meta_h, meta_d, data_d, bound, points_d,
                                 in_mask_d, 1, n, nullptr, ws.cand, ws.passf,
                                 st);
  if (rc) return rc;
  if (which == 1 || rec.kind() == 0) {
    k_fill_u8<<<blocks_for(n, 256), 256, 0, st>>>(ws.passf, n, 1);
    NB_LAUNCH_OK();
  } else {
    rc = neural_any(meta_h, meta_d, data_d, bound, points_d, ws.cand, 1, n, ws,
                    mlp_mode, st);
    if (rc) return rc;
  }
  return launch_apply(3, n, nullptr, ws.cand, ws.passf, out_d, nullptr, st);
}

int nb200_stats(const double* log_l_d, const uint8_t* code_d, int64_t n,
                double log_l_min, double* lse_d, int64_t* counters_d,
                void* workspace_d, size_t workspace_bytes, void* stream) {
  NB_CHECK(n >= 0, "negative n");
  NB_CHECK(workspace_bytes >= sizeof(StatPartial) * STAT_MAX_BLOCKS,
           "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  StatPartial* partial = (StatPartial*)workspace_d;
  int64_t nb = (n + 4 * STAT_THREADS - 1) / (4 * STAT_THREADS);
  if (nb < 1) nb = 1;
  if (nb > STAT_MAX_BLOCKS) nb = STAT_MAX_BLOCKS;
  ProfScope prof(ST_STATS, st);
  k_stats_partial<<<(unsigned)nb, STAT_THREADS, 0, st>>>(
      log_l_d, code_d, n, log_l_min, partial);
  NB_LAUNCH_OK();
  k_stats_final<<<1, STAT_THREADS, 0, st>>>(partial, (int)nb, lse_d,
                                            (long long*)counters_d);
  NB_LAUNCH_OK();
  return 0;
}

int nb200_loglike(const double* points_d, const uint8_t* code_d, int64_t n,
                  int d, int like_id, const double* params_d, int n_params,
                  double* log_l_d, void* stream) {
  NB_CHECK(d >= 1 && d <= NB200_D_MAX, "n_dim out of range");
  NB_CHECK(like_id >= 0 && like_id <= 3, "unknown like_id");
  NB_CHECK(n >= 0, "negative n");
  const int need = like_id == NB200_LIKE_GAUSSIAN ? 2 + d
                   : like_id == NB200_LIKE_ROSENBROCK ? 2
                   : like_id == NB200_LIKE_MIXTURE ? 3 + d
                                                   : 3 + d;
  NB_CHECK(n_params >= need, "too few likelihood parameters");
  if (n == 0) return 0;
  const int threads = threads_for(d);
  const size_t smem = smem_rows(threads, d, 1);
  if (opt_in_smem(k_loglike, smem)) return 2;
  ProfScope prof(ST_LOGLIKE, (cudaStream_t)stream);
  k_loglike<<<blocks_for(n, threads), threads, smem, (cudaStream_t)stream>>>(
      points_d, code_d, n, d, like_id, params_d, log_l_d);
  NB_LAUNCH_OK();
  return 0;
}

int nb200_compact(const double* points_d, const double* log_l_d,
                  const uint8_t* code_d, int64_t n, int d,
                  double* out_points_d, double* out_log_l_d, int64_t* n_out_d,
                  void* workspace_d, size_t workspace_bytes, void* stream) {
  NB_CHECK(n >= 0, "negative n");
  NB_CHECK(workspace_bytes >= nb200_workspace_bytes(n, d),
           "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws;
  workspace_layout(n < 1 ? 1 : n, d, (char*)workspace_d, &ws);
  const int nblocks = (int)((n + CMP_ITEMS - 1) / CMP_ITEMS);
  if (nblocks == 0) {
    NB_CUDA(cudaMemsetAsync(n_out_d, 0, sizeof(int64_t), st));
    return 0;
  }
  ProfScope prof(ST_COMPACT, st);
  k_compact_count<<<nblocks, CMP_THREADS, 0, st>>>(code_d, n, ws.block_count);
  NB_LAUNCH_OK();
  k_compact_scan<<<1, 1024, 0, st>>>(ws.block_count, nblocks,
                                     (long long*)n_out_d);
  NB_LAUNCH_OK();
  if (out_points_d || out_log_l_d) {
    k_compact_scatter<<<nblocks, CMP_THREADS, 0, st>>>(
        points_d, log_l_d, code_d, n, d, ws.block_count, out_points_d,
        out_log_l_d);
    NB_LAUNCH_OK();
  }
  return 0;
}

int nb200_compact_index(const double* log_l_d, const uint8_t* code_d,
                        int64_t n, uint64_t offset, uint64_t* out_index_d,
                        double* out_log_l_d, int64_t* n_out_d,
                        void* workspace_d, size_t workspace_bytes,
                        void* stream) {
  NB_CHECK(n >= 0, "negative n");
  NB_CHECK(code_d && out_index_d && n_out_d, "null argument");
  NB_CHECK(workspace_bytes >= nb200_workspace_bytes(n, 1),
           "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws;
  workspace_layout(n < 1 ? 1 : n, 1, (char*)workspace_d, &ws);
  const int nblocks = (int)((n + CMP_ITEMS - 1) / CMP_ITEMS);
  if (nblocks == 0) {
    NB_CUDA(cudaMemsetAsync(n_out_d, 0, sizeof(int64_t), st));
    return 0;
  }
  ProfScope prof(ST_COMPACT, st);
  k_compact_count<<<nblocks, CMP_THREADS, 0, st>>>(code_d, n, ws.block_count);
  NB_LAUNCH_OK();
  k_compact_index<<<nblocks, CMP_THREADS, 0, st>>>(
      log_l_d, code_d, n, (unsigned long long)offset, ws.block_count,
      (unsigned long long*)out_index_d, out_log_l_d, (long long*)n_out_d);
  NB_LAUNCH_OK();
  return 0;
}

int nb200_materialize(const int32_t* meta_h, const int32_t* meta_d,
                      const double* data_d, int bound, uint64_t seed,
                      uint32_t stream_id, int mlp_mode,
                      const uint64_t* index_d, int64_t k, double* points_out_d,
                      void* stream) {
  if (check_bound(meta_h, bound)) return 1;
  NB_CHECK(k >= 0, "negative k");
  if (k == 0) return 0;
  NB_CHECK(index_d && points_out_d, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned long long* gather = (const unsigned long long*)index_d;
  // the SAME kernel the cycle would run for this bound and emulator
  // arithmetic (their sums round differently)
  if (mode_tc(mlp_mode) && front_applicable(meta_h, bound, nullptr, nullptr))
    return launch_front(meta_h, meta_d, data_d, bound, k, seed, 0, stream_id,
                        points_out_d, nullptr, nullptr, nullptr, -1, nullptr,
                        nullptr, gather, 0, st);
  const Rec rec = record(meta_h, bound);
  const int d = rec.d();
  const int threads = threads_for(d);
  const size_t smem = smem_rows(threads, d, 2);
  if (opt_in_smem(k_union_propose<false>, smem)) return 2;
  ProfScope prof(ST_PROPOSE, st);
  k_union_propose<false><<<blocks_for(k, threads), threads, smem, st>>>(
      meta_d, (int)(rec.r - meta_h), data_d, k, seed, 0, stream_id, nullptr,
      nullptr, nullptr, nullptr, nullptr, points_out_d, nullptr, nullptr,
      gather);
  NB_LAUNCH_OK();
  return 0;
}

int nb200_cycle(const int32_t* meta_h, const int32_t* meta_d,
                const double* data_d, int bound, int first_later, int n_later,
                int64_t n, uint64_t seed, uint64_t offset, uint32_t stream_id,
                int like_id, const double* like_params_d, int n_like_params,
                double log_l_min, int mlp_mode, double* points_d,
                double* log_l_d, uint8_t* code_d, double* lse_d,
                int64_t* counters_d, void* workspace_d, size_t workspace_bytes,
                void* stream) {
  if (check_bound(meta_h, bound)) return 1;
  NB_CHECK(n >= 0, "negative n");
  NB_CHECK(n_later >= 0 && (n_later == 0 || (first_later >= 0 &&
           first_later + n_later <= meta_h[0])), "later-bound range");
  cudaStream_t st = (cudaStream_t)stream;
  const Rec rec = record(meta_h, bound);
  const int d = rec.d();
  NB_CHECK(workspace_bytes >= nb200_workspace_bytes(n, d),
           "workspace too small");
  Workspace ws;
  workspace_layout(n < 1 ? 1 : n, d, (char*)workspace_d, &ws);
  int rc;
  const bool fused = mode_tc(mlp_mode) && n > 0 &&
                     front_applicable(meta_h, bound, nullptr, nullptr);
  if (fused && mlp_mode == NB200_MLP_F16)
    NB_CHECK(mlp_tf32_resident(meta_h, bound, 0, NB200_MLP_F16),
             "this emulator has no fp16 tensor-core blob; use NB200_MLP_TF32");
  bool fused_tail = false;
  int n_partial = 0;
  if (fused) {
    // without later bounds the likelihood and the shell sums ride along: the
    // front kernel evaluates the likelihood of every row that is still a
    // candidate (it holds the row on chip anyway), the emulator kernel sums
    // the survivors
    fused_tail = n_later == 0 && like_id >= 0 && lse_d && counters_d &&
                 mlp_tf32_resident(meta_h, bound, 0, mlp_mode);
    // 1+2 fused: proposal, cube cut, overlap acceptance, neural-ellipsoid
    // whitening and standardisation in one fp64 kernel, then the emulator on
    // tensor cores writes NN rejects straight into the disposition bytes
    rc = launch_front(meta_h, meta_d, data_d, bound, n, seed, offset,
                      stream_id, points_d, code_d, ws.maskj, ws.xs32,
                      fused_tail ? like_id : -1, like_params_d,
                      fused_tail ? log_l_d : nullptr, nullptr,
                      mlp_mode == NB200_MLP_F16, st);
    if (rc) return rc;
    {
      ProfScope prof(ST_MLP, st);
      rc = launch_mlp_tf32_rows(
          meta_h, data_d, bound, 0, ws.xs32, ws.maskj, n, code_d, log_l_min,
          log_l_d, fused_tail ? (void*)ws.partial : nullptr, &n_partial,
          mlp_mode, st);
    }
    if (rc) return rc;
    if (fused_tail) {
      NB_CHECK(n_partial > 0, "resident emulator kernel expected");
      ProfScope prof(ST_STATS, st);
      k_stats_final<<<1, STAT_THREADS, 0, st>>>(ws.partial, n_partial, lse_d,
                                                (long long*)counters_d);
      NB_LAUNCH_OK();
      return 0;
    }
  } else {
    // 1. raw draws, cube filter, overlap acceptance
    rc = nb200_union_propose(meta_h, meta_d, data_d, bound, n, seed, offset,
                             stream_id, nullptr, nullptr, nullptr, nullptr,
                             nullptr, points_d, code_d, nullptr, stream);
    if (rc) return rc;
  }
  if (n > 0) {
    // 2. neural filter of NautilusBound.sample
    if (!fused && rec.kind() == 1 && rec.J() > 0) {
      rc = launch_apply(0, n, code_d, ws.cand, ws.passf, nullptr, nullptr, st);
      if (rc) return rc;
      rc = neural_any(meta_h, meta_d, data_d, bound, points_d, ws.cand, 1, n,
                      ws, mlp_mode, st);
      if (rc) return rc;
      rc = launch_apply(1, n, code_d, ws.cand, ws.passf, nullptr, nullptr, st);
      if (rc) return rc;
    }
    // 3. exclusion by every later bound (sampler.py:796-801): one grouped
    // pass over all (later bound, neural bound) pairs when the emulators run
    // on the tensor cores and the scratch holds at least one pass ...
    bool grouped = false;
    int n_pairs = 0;
    const int32_t* hdr = nullptr;
    if (n_later > 0 && mode_tc(mlp_mode) &&
        excl_applicable(meta_h, first_later, n_later, d, mlp_mode, &n_pairs,
                        &hdr)) {
      const size_t base_bytes = nb200_workspace_bytes(n, d);
      const bool f16 = mlp_mode == NB200_MLP_F16;
      // (scratch rows are sized for the tf32 form; the fp16 rows are shorter)
      const int k0p = k0p_of(d);
      const int k0p_rows = f16 ? (d + 1 + 15) / 16 * 16 : k0p;
      ExclWs ew;
      excl_layout(n, n_pairs, n_later, k0p, (char*)workspace_d + base_bytes,
                  workspace_bytes - base_bytes, -1, &ew);
      if (ew.cap >= 1024 || ew.cap * 1 >= (n + 127) / 128 * 128) {
        grouped = true;
        ProfScope prof(ST_UNION, st);
        rc = launch_excl_pairs(meta_d, first_later, n_later, f16, ew.pairs,
                               ew.pair_base, st);
        if (rc) return rc;
        // candidates = proposals still in the shell
        const int nblocks = (int)((n + CMP_ITEMS - 1) / CMP_ITEMS);
        k_compact_count<<<nblocks, CMP_THREADS, 0, st>>>(code_d, n,
                                                         ws.block_count);
        NB_LAUNCH_OK();
        k_compact_index<<<nblocks, CMP_THREADS, 0, st>>>(
            nullptr, code_d, n, 0ull, ws.block_count, ew.cand_idx, nullptr,
            ws.total);
        NB_LAUNCH_OK();
        NB_CUDA(cudaMemsetAsync(ew.excl, 0, (size_t)n, st));
        const unsigned long long* n_cand =
            (const unsigned long long*)ws.total;
        for (int64_t lo = 0; lo < n; lo += ew.cap) {
          // (passes beyond the candidate count return at once: the count
          // lives on the device, the number of passes must not)
          NB_CUDA(cudaMemsetAsync(ew.seg_count, 0,
                                  sizeof(unsigned int) * (size_t)(n_pairs > 0 ? n_pairs : 1), st));
          ProfScope prof_prep(ST_PREP, st);
          rc = launch_excl_prep(meta_d, data_d, first_later, n_later, d,
                                k0p_rows, f16, points_d, ew.cand_idx, n_cand,
                                lo, ew.cap,
                                ew.pair_base, ew.seg_count, ew.xs, ew.cid,
                                ew.cap, ew.excl, st);
          if (rc) return rc;
          if (hdr) {
            TcGroupArgs G;
            G.n_pairs = n_pairs; G.chunk_lo = lo; G.seg_stride = ew.cap;
            G.pairs = ew.pairs; G.seg_count = ew.seg_count; G.n_cand = n_cand;
            G.cid = ew.cid; G.data = data_d; G.excl = ew.excl;
            ProfScope prof_mlp(ST_GLUE, st);
            rc = run_mlp_tf32_grouped(hdr, ew.xs, G, st);
            if (rc) return rc;
          }
        }
        rc = launch_excl_apply(ew.cand_idx, n_cand, ew.excl, code_d, n, st);
        if (rc) return rc;
      }
    }
    // ... else one bound at a time (no short-circuit, like the reference)
    for (int l = first_later; !grouped && l < first_later + n_later; ++l) {
      if (check_bound(meta_h, l)) return 1;
      NB_CHECK(record(meta_h, l).d() == d, "later bound has another n_dim");
      rc = union_contains_launch(meta_h, meta_d, data_d, l, points_d, code_d,
                                 NB200_CODE_IN_SHELL, n, nullptr, ws.cand,
                                 ws.passf, st);
      if (rc) return rc;
      rc = neural_any(meta_h, meta_d, data_d, l, points_d, ws.cand, 1, n, ws,
                      mlp_mode, st);
      if (rc) return rc;
      rc = launch_apply(2, n, code_d, ws.cand, ws.passf, nullptr, nullptr, st);
      if (rc) return rc;
    }
    // 4. likelihood
    if (like_id >= 0) {
      rc = nb200_loglike(points_d, code_d, n, d, like_id, like_params_d,
                         n_like_params, log_l_d, stream);
      if (rc) return rc;
    }
  }
  // 5. importance-weight sums and counters
  if (lse_d && counters_d) {
    rc = nb200_stats(like_id >= 0 ? log_l_d : nullptr, code_d, n, log_l_min,
                     lse_d, counters_d, ws.partial,
                     sizeof(StatPartial) * STAT_MAX_BLOCKS, stream);
    if (rc) return rc;
  }
  return 0;
}

int nb200_cycle_host(const int32_t* meta_h, int64_t n_meta,
                     const double* data_h, int64_t n_data, int bound,
                     int first_later, int n_later, int64_t n, uint64_t seed,
                     uint64_t offset, uint32_t stream_id, int like_id,
                     const double* like_params_h, int n_like_params,
                     double log_l_min, int mlp_mode, int64_t cap,
                     double* points_out_h, double* log_l_out_h,
                     int64_t* n_out_h, double* lse_h, int64_t* counters_h) {
  if (check_bound(meta_h, bound)) return 1;
  NB_CHECK(n >= 0 && cap >= 0, "negative size");
  const int d = record(meta_h, bound).d();
  const size_t wsb = nb200_workspace_bytes(n, d);
  int32_t* meta_d = nullptr;
  double *data_d = nullptr, *like_d = nullptr, *points = nullptr,
         *log_l = nullptr, *out_points = nullptr, *out_log_l = nullptr,
         *lse = nullptr;
  uint8_t* code = nullptr;
  int64_t* counters = nullptr;
  void* ws = nullptr;
  int rc = 0;
  const size_t nn = (size_t)(n < 1 ? 1 : n);
  auto cleanup = [&]() {
    cudaFree(meta_d); cudaFree(data_d); cudaFree(like_d); cudaFree(points);
    cudaFree(log_l); cudaFree(out_points); cudaFree(out_log_l); cudaFree(lse);
    cudaFree(code); cudaFree(counters); cudaFree(ws);
  };
#define NB_TRY(expr)                         \
  do {                                       \
    cudaError_t e_ = (expr);                 \
    if (e_ != cudaSuccess) {                 \
      snprintf(g_err, sizeof(g_err), "nautilus_b200: CUDA error '%s' at %s:%d", \
               cudaGetErrorString(e_), __FILE__, __LINE__);                \
      cleanup();                             \
      return 2;                              \
    }                                        \
  } while (0)
  NB_TRY(cudaMalloc(&meta_d, sizeof(int32_t) * n_meta));
  NB_TRY(cudaMalloc(&data_d, sizeof(double) * n_data));
  NB_TRY(cudaMalloc(&like_d, sizeof(double) * (n_like_params + 1)));
  NB_TRY(cudaMalloc(&points, sizeof(double) * nn * d));
  NB_TRY(cudaMalloc(&log_l, sizeof(double) * nn));
  NB_TRY(cudaMalloc(&out_points, sizeof(double) * nn * d));
  NB_TRY(cudaMalloc(&out_log_l, sizeof(double) * nn));
  NB_TRY(cudaMalloc(&lse, sizeof(double) * NB200_N_LSE));
  NB_TRY(cudaMalloc(&code, nn));
  NB_TRY(cudaMalloc(&counters, sizeof(int64_t) * (NB200_N_CNT + 1)));
  NB_TRY(cudaMalloc(&ws, wsb));
  NB_TRY(cudaMemcpy(meta_d, meta_h, sizeof(int32_t) * n_meta,
                    cudaMemcpyHostToDevice));
  NB_TRY(cudaMemcpy(data_d, data_h, sizeof(double) * n_data,
                    cudaMemcpyHostToDevice));
  if (n_like_params > 0)
    NB_TRY(cudaMemcpy(like_d, like_params_h, sizeof(double) * n_like_params,
                      cudaMemcpyHostToDevice));
  rc = nb200_cycle(meta_h, meta_d, data_d, bound, first_later, n_later, n,
                   seed, offset, stream_id, like_id, like_d, n_like_params,
                   log_l_min, mlp_mode, points, log_l, code, lse, counters, ws,
                   wsb, nullptr);
  if (!rc)
    rc = nb200_compact(points, log_l, code, n, d, out_points, out_log_l,
                       counters + NB200_N_CNT, ws, wsb, nullptr);
  if (rc) { cleanup(); return rc; }
  int64_t cnt[NB200_N_CNT + 1];
  NB_TRY(cudaMemcpy(cnt, counters, sizeof(cnt), cudaMemcpyDeviceToHost));
  const int64_t n_out = cnt[NB200_N_CNT];
  if (n_out > cap) {
    cleanup();
    return fail("nautilus_b200: %s (need %lld rows, cap %lld)",
                "output capacity too small", n_out, cap);
  }
  if (n_out > 0) {
    NB_TRY(cudaMemcpy(points_out_h, out_points, sizeof(double) * n_out * d,
                      cudaMemcpyDeviceToHost));
    NB_TRY(cudaMemcpy(log_l_out_h, out_log_l, sizeof(double) * n_out,
                      cudaMemcpyDeviceToHost));
  }
  NB_TRY(cudaMemcpy(lse_h, lse, sizeof(double) * NB200_N_LSE,
                    cudaMemcpyDeviceToHost));
  memcpy(counters_h, cnt, sizeof(int64_t) * NB200_N_CNT);
  *n_out_h = n_out;
  cleanup();
#undef NB_TRY
  return 0;
}

}  // extern "C"
