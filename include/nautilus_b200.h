/*
 * nautilus_b200 -- C ABI of the B200-native importance-nested-sampling cycle.
 *
 * The reference (johannesulf/nautilus v1.0.6) is pure Python and has no FFI;
 * its seam is Python duck typing.  This header is the boundary
 * a maintainer would bind (ctypes): every entry point is
 * `extern "C"`, takes plain pointers / sizes / scalars, and names the reference
 * method it replaces.  All `file:line` citations are relative to
 * /root/reference/nautilus/.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure;
 *     nb200_last_error() returns a thread-local message for the last failure.
 *   - `*_d` pointers are DEVICE pointers on the current CUDA device, `*_h`
 *     pointers are HOST pointers.  `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream).  Device-pointer entry points only
 *     enqueue work; they do not synchronise.
 *   - points are row-major float64 [n, d]; masks / flags are uint8; counters
 *     are int64.  d <= NB200_D_MAX.
 *   - bounds are passed as a serialised *stack* (meta int32[], data float64[])
 *     produced by nautilus_b200/_pack.py:pack_stack; `bound` selects a record.
 *     The host copy of `meta` is needed because launch shapes depend on it.
 *
 * Blob layout (mirrored by nautilus_b200/csrc/nb200_common.cuh)
 *   meta[0] = L (#bounds); meta[1+i] = start of record i.
 *   record header (16 ints): len, kind(0 cube,1 nautilus), d, K, J, unit,
 *       off_cdf(data), off_mix(rel), off_neural(rel), max_width,
 *       same_k (1 + index of the mixture whose ellipsoid IS neural bound 0's,
 *       0: none), amp_log2 (with same_k: ceil(log2(|B_inv|_2 (|B|_2 +
 *       max|c|))) of that ellipsoid -- how much the round trip x = c + B z ->
 *       B_inv (x - c) can amplify rounding; bounds the guard band of the
 *       fused front end's whitening shortcut), 0...
 *   mixture record (8 ints): de, nc, off_idx(rel; d ints: ellipsoid dims then
 *       cube dims), off_c, off_B, off_Binv (data; -1 if de==0), binv_is_lower, 0
 *   neural record (12 ints): off_c, off_Binv, binv_is_lower, n_net, n_lay,
 *       off_mean, off_scale, off_thr(data: {score_predict_min-1e-9, raw}),
 *       off_sizes(rel; n_lay+1 ints), off_wtab(rel; n_net*n_lay*{off_W,off_b}),
 *       off_tc(data; tf32 weight blob for tcgen05, -1 if none),
 *       off_tc_hdr(rel; 32 ints, TcHeader in csrc/nb200_mlp_tc.cu)
 */
#ifndef NAUTILUS_B200_H
#define NAUTILUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB200_VERSION 200
#define NB200_D_MAX 128
#define NB200_W_MAX 256

/* disposition of a raw proposal (one byte per proposal) */
#define NB200_CODE_CUBE_REJECT 0     /* bounds/union.py:313-314            */
#define NB200_CODE_OVERLAP_REJECT 1  /* bounds/union.py:316-319            */
#define NB200_CODE_NN_REJECT 2       /* bounds/nautilus.py:217-219         */
#define NB200_CODE_EXCLUDED 3        /* sampler.py:796-801                 */
#define NB200_CODE_IN_SHELL 4        /* likelihood evaluated               */

/* emulator arithmetic */
#define NB200_MLP_F64 0   /* fp64 CUDA cores, canonical summation order     */
#define NB200_MLP_TF32 1  /* tcgen05 kind::tf32, fp32 accumulate in TMEM    */
#define NB200_MLP_F16 2   /* tcgen05 kind::f16 (fp16 operands: the same 10
                           * mantissa bits as tf32, half the bytes), fp32
                           * accumulate; emulators with resident weights    */

/* built-in synthetic likelihoods; params are float64        */
#define NB200_LIKE_GAUSSIAN 0    /* {inv_sigma2, norm, mu[d]}                */
#define NB200_LIKE_ROSENBROCK 1  /* {lo, width}  theta = lo + width * u      */
#define NB200_LIKE_MIXTURE 2     /* {M, inv_sigma2, norm, mu[M*d]}           */
#define NB200_LIKE_EQUICORR 3    /* {a, b, norm, mu[d]}: -0.5(a*S2 - b*S1^2) */

/* indices into the int64 counters written by nb200_cycle / nb200_stats      */
#define NB200_CNT_RAW 0
#define NB200_CNT_CUBE_REJECT 1
#define NB200_CNT_OVERLAP_REJECT 2
#define NB200_CNT_NN_REJECT 3
#define NB200_CNT_EXCLUDED 4
#define NB200_CNT_IN_SHELL 5
#define NB200_CNT_UPDATE 6 /* #(log_l >= log_l_min), sampler.py:1144        */
#define NB200_N_CNT 8
/* float64 sums: {max, sum exp(l-max), sum exp(2(l-max)), 0}                 */
#define NB200_N_LSE 4

const char* nb200_last_error(void);
int nb200_version(void);
/* sm_count, compute capability of the current device */
int nb200_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* number of kernels this library has launched in this process (for
 * bench.py's gpu_launches) */
int64_t nb200_launch_count(void);

/* Optional per-stage CUDA-event timing (used by bench.py for the roofline of
 * the dominant kernel; adds two event records per stage, keep it off when
 * measuring throughput).  collect() synchronises the device and returns the
 * accumulated milliseconds and scope counts of NB200_N_STAGES stages. */
#define NB200_N_STAGES 10
void nb200_profile_enable(int on);
int nb200_profile_collect(double* ms_by_stage, int64_t* calls_by_stage);
const char* nb200_profile_stage_name(int stage);

/* ---- Ellipsoid (bounds/basic.py:244-449) ------------------------------ */

/* Ellipsoid.transform, basic.py:339-342.  inverse==0: out = M (x - c) with
 * M = B_inv; inverse!=0: out = M x + c with M = B.  M is dense d*d row-major. */
int nb200_ell_transform(const double* points_d, int64_t n, int d,
                        const double* c_d, const double* M_d, int inverse,
                        double* out_d, void* stream);
/* Ellipsoid.contains, basic.py:360: sum((B_inv (x-c))^2) < 1 (strict).
 * r2_d may be NULL; otherwise receives the squared Mahalanobis radius. */
int nb200_ell_contains(const double* points_d, int64_t n, int d,
                       const double* c_d, const double* Binv_d,
                       uint8_t* out_d, double* r2_d, void* stream);
/* Ellipsoid.sample with explicit base randoms, basic.py:376-381:
 * z f64[n,d] normals, u f64[n] uniforms. */
int nb200_ell_sample_from(const double* z_d, const double* u_d, int64_t n,
                          int d, const double* c_d, const double* B_d,
                          double* out_d, void* stream);

/* ---- Union / UnitCubeEllipsoidMixture (bounds/union.py, basic.py:452) -- */

/* Union.contains (union.py:285-289) and the overlap count
 * n_bound = sum_k contains_k (union.py:316-317).  in_mask_d may be NULL (all
 * points).  count_d / contains_d may each be NULL. */
int nb200_union_count(const int32_t* meta_h, const int32_t* meta_d,
                      const double* data_d, int bound,
                      const double* points_d, const uint8_t* in_mask_d,
                      int64_t n, int32_t* count_d, uint8_t* contains_d,
                      void* stream);

/* Raw draws of one Union.sample pass (union.py:305-319) for proposals
 * offset .. offset+n-1: choose ellipsoid k ~ p_k, draw inside mixture k,
 * unit-cube filter, overlap count, accept iff r > 1 - 1/n_bound.
 * Philox mode: k_d == NULL; randoms come from Philox4x32-10 keyed by `seed`
 * with counter (proposal index, block, stream_id).
 * Test mode: k_d i32[n], z_d f64[n,d] (first de columns used), cube_u_d
 * f64[n,d] (first nc columns), u_d f64[n], r_d f64[n] supplied by the host
 * exactly as the reference drew them.
 * Outputs: points f64[n,d], code u8[n] (0, 1 or 4), n_bound i32[n] (may be
 * NULL).  For a UnitCube record (sampler shell 0, basic.py:85) points are
 * uniform and code is 4. */
int nb200_union_propose(const int32_t* meta_h, const int32_t* meta_d,
                        const double* data_d, int bound, int64_t n,
                        uint64_t seed, uint64_t offset, uint32_t stream_id,
                        const int32_t* k_d, const double* z_d,
                        const double* cube_u_d, const double* u_d,
                        const double* r_d, double* points_d, uint8_t* code_d,
                        int32_t* n_bound_d, void* stream);

/* ---- NeuralNetworkEmulator (neural.py:100-116) ------------------------- */

/* emulator.predict of neural bound j of `bound` on whitened coordinates
 * x f64[n,d] (the standardisation (x-mean)/scale is applied inside). */
int nb200_mlp_predict(const int32_t* meta_h, const int32_t* meta_d,
                      const double* data_d, int bound, int j,
                      const double* x_d, int64_t n, double* out_d,
                      int mlp_mode, void* workspace_d, size_t workspace_bytes,
                      void* stream);

/* NeuralNetworkEmulator.train (neural.py:50-98): fit n_net networks with
 * layer sizes sizes_h[0..n_lay] (sizes_h[0] == d, sizes_h[n_lay] == 1) to the
 * already standardised inputs x f64[m,d] and targets y f64[m] with sklearn
 * MLPRegressor's algorithm (Adam, squared loss, ReLU, minibatch, patience
 * stopping; see csrc/nb200_mlp_fit.cu).  Network i is seeded with (seed, i),
 * like random_state=i in the reference (neural.py:94).  Outputs per network:
 * parameters f64[n_params] in the order W_0 (fan_in x fan_out, row-major),
 * b_0, W_1, b_1, ...; epochs run; final epoch loss. */
size_t nb200_mlp_fit_workspace_bytes(int64_t m, int d, int n_params,
                                     int n_net);
int nb200_mlp_fit(const double* x_d, const double* y_d, int64_t m, int d,
                  const int32_t* sizes_h, int n_lay, int n_net, uint64_t seed,
                  double lr, double beta1, double beta2, double eps,
                  int batch_size, int max_epochs, double tol, int patience,
                  double* weights_out_d, int32_t* n_iter_out_d,
                  double* loss_out_d, void* workspace_d,
                  size_t workspace_bytes, void* stream);

/* ---- NautilusBound.contains (bounds/nautilus.py:146-169) --------------- */

/* which: 0 = full bound (union & any neural), 1 = union only,
 * 2 = any NeuralBound only (the filter of NautilusBound.sample,
 * nautilus.py:217-218).  in_mask_d may be NULL. */
int nb200_bound_contains(const int32_t* meta_h, const int32_t* meta_d,
                         const double* data_d, int bound, int which,
                         const double* points_d, const uint8_t* in_mask_d,
                         int64_t n, uint8_t* out_d, int mlp_mode,
                         void* workspace_d, size_t workspace_bytes,
                         void* stream);

/* bytes of scratch the bound-level entry points need for n points in d dims */
size_t nb200_workspace_bytes(int64_t n, int d);
/* ... and what nb200_cycle would like when `n_pairs` (later bound, neural
 * bound) pairs exclude proposals (sum of J over the later bounds): with this
 * much the exclusion of sampler.py:796-801 runs as ONE grouped pass per 2^18
 * candidates whatever the number of later bounds; with less (but at least
 * nb200_workspace_bytes) it runs in more passes, or one bound at a time. */
size_t nb200_cycle_workspace_bytes(int64_t n, int d, int n_pairs);

/* ---- shell reductions (sampler.py:925-943, 1144) ----------------------- */

/* One-pass (max, sum e^{l-m}, sum e^{2(l-m)}) over log_l[i] with code[i]==4
 * (code_d NULL = all), plus the disposition histogram and
 * #(log_l >= log_l_min).  lse_d f64[4], counters_d i64[8].  Deterministic:
 * fixed-shape tree, no atomics. */
int nb200_stats(const double* log_l_d, const uint8_t* code_d, int64_t n,
                double log_l_min, double* lse_d, int64_t* counters_d,
                void* workspace_d, size_t workspace_bytes, void* stream);

/* built-in likelihood on points with code==4 (code_d NULL = all) */
int nb200_loglike(const double* points_d, const uint8_t* code_d, int64_t n,
                  int d, int like_id, const double* params_d, int n_params,
                  double* log_l_d, void* stream);

/* stable compaction of the rows with code==4 into out_points / out_log_l
 * (either may be NULL); *n_out_d receives the count. */
int nb200_compact(const double* points_d, const double* log_l_d,
                  const uint8_t* code_d, int64_t n, int d,
                  double* out_points_d, double* out_log_l_d, int64_t* n_out_d,
                  void* workspace_d, size_t workspace_bytes, void* stream);

/* Index form of the compaction: for the rows with code==4 the GLOBAL
 * proposal index offset + i (u64: the Philox counter of the proposal, which
 * nb200_materialize turns back into the row) and log_l (may be NULL).  What
 * Sampler.add_samples keeps of a batch (sampler.py:1135-1141) is then 16
 * bytes per point instead of 8 d + 8; the rows stay on the device. */
int nb200_compact_index(const double* log_l_d, const uint8_t* code_d,
                        int64_t n, uint64_t offset, uint64_t* out_index_d,
                        double* out_log_l_d, int64_t* n_out_d,
                        void* workspace_d, size_t workspace_bytes,
                        void* stream);

/* Rows of the proposals with global indices index_d[0..k) of stack record
 * `bound` (a proposal is a pure function of (bound, seed, stream_id, index)):
 * re-runs the proposal kernel the cycle used for this bound and `mlp_mode`
 * on exactly these indices; points_out_d f64[k,d] is BIT-IDENTICAL to the
 * rows nb200_cycle wrote for them (tests/test_gpu_session.py).  Replaces the
 * `points[in_shell]` copies of sampler.py:801,1135 by an on-demand gather. */
int nb200_materialize(const int32_t* meta_h, const int32_t* meta_d,
                      const double* data_d, int bound, uint64_t seed,
                      uint32_t stream_id, int mlp_mode,
                      const uint64_t* index_d, int64_t k, double* points_out_d,
                      void* stream);

/* ---- the full cycle (sampler.py:1093-1144 over one raw batch) ---------- */

/* n raw proposals from stack record `bound`, filtered by its neural bounds,
 * excluded by records first_later .. first_later+n_later-1, likelihood
 * `like_id` (or none if like_id < 0), reductions.  Outputs: points f64[n,d],
 * log_l f64[n] (NaN unless code 4), code u8[n], lse f64[4], counters i64[8].
 */
int nb200_cycle(const int32_t* meta_h, const int32_t* meta_d,
                const double* data_d, int bound, int first_later, int n_later,
                int64_t n, uint64_t seed, uint64_t offset, uint32_t stream_id,
                int like_id, const double* like_params_d, int n_like_params,
                double log_l_min, int mlp_mode, double* points_d,
                double* log_l_d, uint8_t* code_d, double* lse_d,
                int64_t* counters_d, void* workspace_d, size_t workspace_bytes,
                void* stream);

/* Host-buffer form of the cycle (what a ctypes/NumPy caller binds): uploads
 * the blob, runs nb200_cycle, compacts, and copies the in-shell points and
 * log_l back.  points_out_h f64[cap,d], log_l_out_h f64[cap]; *n_out_h <= cap
 * rows are written (error if more would be needed). */
int nb200_cycle_host(const int32_t* meta_h, int64_t n_meta,
                     const double* data_h, int64_t n_data, int bound,
                     int first_later, int n_later, int64_t n, uint64_t seed,
                     uint64_t offset, uint32_t stream_id, int like_id,
                     const double* like_params_h, int n_like_params,
                     double log_l_min, int mlp_mode, int64_t cap,
                     double* points_out_h, double* log_l_out_h,
                     int64_t* n_out_h, double* lse_h, int64_t* counters_h);

/* ---- live set (sampler.py:1007-1009, 1160-1164) -------------------------- */

/* The k-th largest of values f64[n] (restricted to mask_d != 0 when mask_d is
 * given; at least k entries must be selected): *thr_d receives it, *greater_d
 * the number of selected values strictly greater.  The live set of the
 * sampler is {log_l > thr} plus k - greater of the ties -- what the reference
 * gets from a full argsort of every stored log_l on each exploration step.
 * MSB radix select, four 16-bit passes; nothing leaves the device. */
size_t nb200_select_workspace_bytes(void);
int nb200_select_kth_largest(const double* values_d, const uint8_t* mask_d,
                             int64_t n, int64_t k, double* thr_d,
                             int64_t* greater_d, void* workspace_d,
                             size_t workspace_bytes, void* stream);

/* ---- bound construction (between shells) -------------------------------- */

/* Weights u f64[n] of Khachiyan's algorithm for the minimum-volume enclosing
 * ellipsoid of n points in d dimensions (minimum_volume_enclosing_ellipsoid,
 * bounds/basic.py:175-241): the ellipsoid is centred at c = sum u_i x_i with
 * shape A^-1 = d * sum u_i (x_i - c)(x_i - c)^T.  qT_d f64[d, n] holds the
 * points COORDINATE-MAJOR (and, for conditioning, whitened by their sample
 * covariance: the MVEE is affine equivariant).  Stops after max_updates
 * rank-one updates or when max_i g_i <= (d+1)(1+tol).  iters_d (may be NULL)
 * receives the number of updates.  workspace: nb200_mvee_workspace_bytes(n). */
size_t nb200_mvee_workspace_bytes(int64_t n);
int nb200_mvee_weights(const double* qT_d, int64_t n, int d, int max_updates,
                       double tol, double* u_d, int32_t* iters_d,
                       void* workspace_d, size_t workspace_bytes,
                       void* stream);

/* Two-component Gaussian mixture by EM, n_init restarts in one launch: the
 * GaussianMixture(n_components=2, n_init=10) call of Union.split
 * (bounds/union.py:185-187).  x_d f64[n, d] row-major (the whitened points of
 * the bound to split); label_d u8[n_init, n] the initial hard assignment of
 * every restart (k-means++ seeding, drawn by the caller's generator).  Every
 * restart runs EM until its mean log likelihood moves by less than tol, at
 * most max_iter steps; reg is added to the diagonal of the covariances.
 * Outputs: log_p_d f64[n_init, 2, n] = log(w_k N(x | mu_k, C_k)) of the last
 * accepted step, score_d f64[n_init] = the restart's mean log likelihood
 * (-inf: abandoned), iters_d i32[n_init].  nb200_gmm2_applicable: 1 if the
 * shape fits the kernel (d up to ~55: both packed moment sets, the factors and a CTA's share of the points or at least
 * their responsibilities in shared memory). */
int nb200_gmm2_applicable(int64_t n, int d);
int nb200_gmm2_em(const double* x_d, int64_t n, int d,
                  const uint8_t* label_d, int n_init, int max_iter, double tol,
                  double reg, double* log_p_d, double* score_d,
                  int32_t* iters_d, void* stream);

/* ---- host-buffer session: the loop around add_samples ------------------- */

/* A session owns the device buffers, two streams and pinned host buffers for
 * batches of up to n_max raw proposals of one bound stack, so that a host
 * caller (NumPy / ctypes, like the reference's Sampler) pays no allocation
 * per batch and the device->host copy of batch i overlaps the kernels of
 * batch i+1.  `cap` = most in-shell rows one batch may return (<= n_max),
 * n_slots (1..4) = batches that may be in flight.  The stack is copied; it
 * can be replaced by set_stack (same or smaller size) between shells
 * (sampler.py:1023-1039, a new bound was accepted).
 *
 *   submit(slot, ...)  enqueues, without blocking the host: [H2D of the stack
 *       if upload_stack] -> H2D of the likelihood parameters -> nb200_cycle ->
 *       nb200_compact -> D2H of sums / counters / row count.
 *   wait(slot, ...)    blocks until that batch is done, copies its in-shell
 *       points and log_l into the slot's pinned buffers and returns pointers
 *       to them (valid until the slot is submitted again): what
 *       Sampler.add_samples appends (sampler.py:1135-1141), plus the inputs
 *       of update_shell_info (sampler.py:925-943) as lse f64[4] and
 *       counters i64[8].  *log_l_h is NULL when like_id < 0.
 * Not thread-safe per session; different sessions are independent. */
typedef struct nb200_session nb200_session;
int nb200_session_create(const int32_t* meta_h, int64_t n_meta,
                         const double* data_h, int64_t n_data, int64_t n_max,
                         int64_t cap, int n_slots, int like_params_max,
                         nb200_session** out);
int nb200_session_destroy(nb200_session* s);
int nb200_session_set_stack(nb200_session* s, const int32_t* meta_h,
                            int64_t n_meta, const double* data_h,
                            int64_t n_data);
int nb200_session_submit(nb200_session* s, int slot, int upload_stack,
                         int bound, int first_later, int n_later, int64_t n,
                         uint64_t seed, uint64_t offset, uint32_t stream_id,
                         int like_id, const double* like_params_h,
                         int n_like_params, double log_l_min, int mlp_mode);
int nb200_session_wait(nb200_session* s, int slot, const double** points_h,
                       const double** log_l_h, int64_t* n_out, double* lse_h,
                       int64_t* counters_h);

/* Index mode: what crosses PCIe per in-shell proposal is (global index u64,
 * log_l f64) = 16 bytes instead of the 8 d + 8 of a row; the rows stay on
 * the device and are regenerated on demand (a row is a pure function of
 * (bound, seed, stream_id, index)).
 *   set_returns(NB200_RETURN_INDEX)  batches submitted from now on compact
 *       indices; their device->host copy is enqueued by submit itself (sized
 *       from the previous batch, topped up by wait if it fell short), so wait
 *       costs ONE host synchronisation;
 *   wait_index(slot, ...)  as wait, with index_h u64[*n_out] instead of rows;
 *   materialize(...)  rows f64[k,d] (host) of any k indices of `bound`, bit-
 *       identical to the rows the cycle produced -- posterior(), add_bound and
 *       the transfer step (sampler.py:541-647, 1007-1089) ask for rows, the
 *       batch loop never does. */
#define NB200_RETURN_ROWS 0
#define NB200_RETURN_INDEX 1
int nb200_session_set_returns(nb200_session* s, int what);
int nb200_session_wait_index(nb200_session* s, int slot,
                             const uint64_t** index_h,
                             const double** log_l_h, int64_t* n_out,
                             double* lse_h, int64_t* counters_h);
int nb200_session_materialize(nb200_session* s, int bound, uint64_t seed,
                              uint32_t stream_id, int mlp_mode,
                              const uint64_t* index_h, int64_t k,
                              double* points_out_h);

#ifdef __cplusplus
}
#endif
#endif /* NAUTILUS_B200_H */
