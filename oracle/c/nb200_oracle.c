/*
 * Canonical-order C restatement of the deterministic kernels of the cycle.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): built by
 * __graft_entry__.build() / oracle/c/Makefile into oracle/c/libnb200_oracle.so
 * and loaded by tests, smoke() and bench.py's cpu_baseline leg -- never by
 * nautilus_b200/.
 *
 * The NumPy oracle (oracle/nautilus_oracle.py) is bit-identical to the
 * reference but inherits NumPy's SIMD / pairwise / BLAS summation orders,
 * which no other implementation can reproduce bit-for-bit.  This file fixes
 * ONE summation order -- left-to-right fused-multiply-add chains, the order
 * the CUDA kernels use (nautilus_b200/csrc/nb200_device.cuh) -- so that the
 * fp64 intermediates (whitened coordinates, squared radii, emulator scores)
 * of the GPU path can be compared bit-for-bit, while booleans / counts are
 * compared against both oracles.  It reads the same serialised bound stack as
 * the kernels (layout: include/nautilus_b200.h).
 *
 * Reference lines restated (relative to /root/reference/nautilus/):
 *   bounds/basic.py:67        UnitCube.contains
 *   bounds/basic.py:339-342   Ellipsoid.transform
 *   bounds/basic.py:360       Ellipsoid.contains
 *   bounds/basic.py:610-617   UnitCubeEllipsoidMixture.contains
 *   bounds/union.py:285-289   Union.contains, :316-317 overlap count
 *   bounds/neural.py:115-126  NeuralBound.contains
 *   bounds/nautilus.py:162-169 NautilusBound.contains
 *   neural.py:114-116         NeuralNetworkEmulator.predict -> scikit-learn
 *       1.9.0 neural_network/_multilayer_perceptron.py:189-224 forward pass
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (no -mfma: fma() from libm
 * is exact and dispatches to the hardware instruction where available).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define HDR 16
#define MIX_REC 8
#define NB_REC 12

static const int32_t* record(const int32_t* meta, int bound) {
  return meta + meta[1 + bound];
}

/* t_i = sum_j M[i][j] s_j, left-to-right FMA chain starting from 0 */
static void matvec(const double* M, int de, const double* s, double* t) {
  for (int i = 0; i < de; ++i) {
    double a = 0.0;
    for (int j = 0; j < de; ++j) a = fma(M[(size_t)i * de + j], s[j], a);
    t[i] = a;
  }
}

static double whiten_r2(const double* x, const int32_t* idx, int de,
                        const double* c, const double* Binv, double* s,
                        double* t) {
  for (int j = 0; j < de; ++j) s[j] = x[idx ? idx[j] : j] - c[j];
  matvec(Binv, de, s, t);
  double r2 = 0.0;
  for (int i = 0; i < de; ++i) r2 = fma(t[i], t[i], r2);
  return r2;
}

static int cube_ok(const double* x, const int32_t* idx, int n) {
  int ok = 1;
  for (int j = 0; j < n; ++j) {
    const double v = x[idx ? idx[j] : j];
    ok = ok && (v >= 0.0) && (v < 1.0);
  }
  return ok;
}

static int mix_contains(const int32_t* rec, const double* data, int k,
                        const double* x, double* s, double* t) {
  const int32_t* m = rec + rec[7] + k * MIX_REC;
  const int de = m[0], nc = m[1];
  const int32_t* idx = rec + m[2];
  int in = 1;
  if (nc > 0) in = cube_ok(x, idx + de, nc);
  if (de > 0) {
    const double r2 = whiten_r2(x, idx, de, data + m[3], data + m[5], s, t);
    in = in && (r2 < 1.0);
  }
  return in;
}

void orc_ell_transform(const double* points, int64_t n, int d, const double* c,
                       const double* M, int inverse, double* out) {
  double* s = (double*)malloc(sizeof(double) * d);
  for (int64_t p = 0; p < n; ++p) {
    const double* x = points + p * d;
    double* o = out + p * d;
    if (!inverse) {
      for (int j = 0; j < d; ++j) s[j] = x[j] - c[j];
      matvec(M, d, s, o);
    } else {
      matvec(M, d, x, o);
      for (int j = 0; j < d; ++j) o[j] = o[j] + c[j];
    }
  }
  free(s);
}

void orc_ell_contains(const double* points, int64_t n, int d, const double* c,
                      const double* Binv, uint8_t* out, double* r2_out) {
  double* s = (double*)malloc(sizeof(double) * d * 2);
  for (int64_t p = 0; p < n; ++p) {
    const double r2 = whiten_r2(points + p * d, NULL, d, c, Binv, s, s + d);
    out[p] = r2 < 1.0;
    if (r2_out) r2_out[p] = r2;
  }
  free(s);
}

void orc_union_count(const int32_t* meta, const double* data, int bound,
                     const double* points, int64_t n, int32_t* count,
                     uint8_t* contains) {
  const int32_t* rec = record(meta, bound);
  const int d = rec[2], K = rec[3];
  double* s = (double*)malloc(sizeof(double) * d * 2);
  for (int64_t p = 0; p < n; ++p) {
    const double* x = points + p * d;
    int cnt = 0, in;
    if (rec[1] == 0) {
      in = cube_ok(x, NULL, d);
      cnt = in;
    } else {
      for (int k = 0; k < K; ++k) cnt += mix_contains(rec, data, k, x, s, s + d);
      in = cnt > 0;
      if (in && rec[5]) in = cube_ok(x, NULL, d);
    }
    if (count) count[p] = cnt;
    if (contains) contains[p] = (uint8_t)in;
  }
  free(s);
}

/* emulator on whitened coordinates t (one point); canonical order:
 * h_n = (sum_k W[k][n] a_k) + b_n, ReLU, ensemble ((p0+p1)+...)/n_net */
static double emulator(const int32_t* rec, const double* data,
                       const int32_t* nb, const double* t, double* xs,
                       double* h0, double* h1) {
  const int d = rec[2], n_net = nb[3], n_lay = nb[4];
  const double* mean = data + nb[5];
  const double* scale = data + nb[6];
  const int32_t* sizes = rec + nb[8];
  const int32_t* wtab = rec + nb[9];
  for (int k = 0; k < d; ++k) xs[k] = (t[k] - mean[k]) / scale[k];
  double sum = 0.0;
  for (int net = 0; net < n_net; ++net) {
    const double* a = xs;
    double* h = h0;
    double* ho = h1;
    double y = 0.0;
    for (int l = 0; l < n_lay; ++l) {
      const int fi = sizes[l], fo = sizes[l + 1];
      const double* W = data + wtab[(net * n_lay + l) * 2];
      const double* b = data + wtab[(net * n_lay + l) * 2 + 1];
      for (int o = 0; o < fo; ++o) {
        double acc = 0.0;
        for (int k = 0; k < fi; ++k) acc = fma(W[(size_t)k * fo + o], a[k], acc);
        double v = acc + b[o];
        if (l != n_lay - 1) {
          v = fmax(v, 0.0);
          h[o] = v;
        } else {
          y = v;
        }
      }
      a = h;
      double* tmp = h; h = ho; ho = tmp;
    }
    sum = net == 0 ? y : sum + y;
  }
  return sum / (double)n_net;
}

/* NeuralBound j of `bound`: in_ell, whitened rows, score (NaN outside), pass */
void orc_neural(const int32_t* meta, const double* data, int bound, int j,
                const double* points, int64_t n, uint8_t* in_ell,
                double* t_rows, double* score, uint8_t* pass) {
  const int32_t* rec = record(meta, bound);
  const int d = rec[2];
  const int32_t* nb = rec + rec[8] + j * NB_REC;
  const int w = rec[9] > d ? rec[9] : d;
  double* buf = (double*)malloc(sizeof(double) * (3 * d + 2 * w));
  double *s = buf, *t = buf + d, *xs = buf + 2 * d, *h0 = buf + 3 * d,
         *h1 = buf + 3 * d + w;
  for (int64_t p = 0; p < n; ++p) {
    const double r2 = whiten_r2(points + p * d, NULL, d, data + nb[0],
                                data + nb[1], s, t);
    const int in = r2 < 1.0;
    if (in_ell) in_ell[p] = (uint8_t)in;
    if (t_rows) for (int k = 0; k < d; ++k) t_rows[p * d + k] = t[k];
    double sc = NAN;
    int ok = in;
    if (in && nb[3] > 0) {
      sc = emulator(rec, data, nb, t, xs, h0, h1);
      ok = sc > data[nb[7]];
    }
    if (score) score[p] = sc;
    if (pass) pass[p] = (uint8_t)ok;
  }
  free(buf);
}

/* emulator.predict on given whitened coordinates */
void orc_mlp_predict(const int32_t* meta, const double* data, int bound, int j,
                     const double* t_rows, int64_t n, double* out) {
  const int32_t* rec = record(meta, bound);
  const int d = rec[2];
  const int32_t* nb = rec + rec[8] + j * NB_REC;
  const int w = rec[9] > d ? rec[9] : d;
  double* buf = (double*)malloc(sizeof(double) * (d + 2 * w));
  for (int64_t p = 0; p < n; ++p)
    out[p] = emulator(rec, data, nb, t_rows + p * d, buf, buf + d, buf + d + w);
  free(buf);
}

/* NautilusBound.contains / UnitCube.contains */
void orc_bound_contains(const int32_t* meta, const double* data, int bound,
                        const double* points, int64_t n, uint8_t* out) {
  const int32_t* rec = record(meta, bound);
  const int J = rec[4];
  orc_union_count(meta, data, bound, points, n, NULL, out);
  if (rec[1] == 0 || J == 0) return;
  uint8_t* pass = (uint8_t*)malloc((size_t)n);
  uint8_t* any = (uint8_t*)calloc((size_t)n, 1);
  for (int j = 0; j < J; ++j) {
    orc_neural(meta, data, bound, j, points, n, NULL, NULL, NULL, pass);
    for (int64_t p = 0; p < n; ++p) any[p] |= pass[p];
  }
  for (int64_t p = 0; p < n; ++p) out[p] = out[p] && any[p];
  free(pass);
  free(any);
}
