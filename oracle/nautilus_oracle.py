"""NumPy restatement of the nautilus hot path (bounds, emulator, shell sums).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  All ``file:line``
citations are relative to ``/root/reference/nautilus/`` (v1.0.6).

Bounds are passed around as plain *spec* dictionaries (the same layout that
``nautilus_b200._pack`` serialises for the device), never as reference
objects, so this module has no dependency on the reference being importable:

    ell  = dict(c=f64[de], B=f64[de,de], B_inv=f64[de,de])
    mix  = dict(dim_cube=bool[d], ell=ell | None)
    emu  = dict(mean=f64[d], scale=f64[d],
                coefs=[[W_0 .. W_L-1] per net], intercepts=[[b_0 ..] per net])
    nb   = dict(ell=ell, emulator=emu | None, score_predict_min=float)
    spec = dict(kind='nautilus', n_dim=d, unit=bool, log_v_all=f64[K],
                mixtures=[mix]*K, neural=[nb]*J)        or
           dict(kind='cube', n_dim=d)
"""

import numpy as np
from scipy.special import gammaln, logsumexp

from . import philox


# --------------------------------------------------------------------------
# UnitCube  (bounds/basic.py:9-151)
# --------------------------------------------------------------------------

def cube_contains(points):
    """``UnitCube.contains``: all(0 <= x < 1)  (bounds/basic.py:67)."""
    return np.all((points >= 0) & (points < 1), axis=-1)


# --------------------------------------------------------------------------
# Ellipsoid  (bounds/basic.py:244-449)
# --------------------------------------------------------------------------

def ell_transform(ell, points, inverse=False):
    """``Ellipsoid.transform`` (bounds/basic.py:339-342).

    forward: t = B_inv (x - c); inverse: x = B z + c, both through the same
    ``einsum('ij, ...j')`` call as the reference so the result is
    bit-identical to it.
    """
    if not inverse:
        return np.einsum('ij, ...j', ell['B_inv'], points - ell['c'])
    return np.einsum('ij, ...j', ell['B'], points) + ell['c']


def ell_contains(ell, points):
    """``Ellipsoid.contains``: sum(t**2) < 1, strict (bounds/basic.py:360)."""
    return np.sum(ell_transform(ell, points)**2, axis=-1) < 1


def ell_sample_from(ell, z, u):
    """``Ellipsoid.sample`` with the base randoms made explicit.

    bounds/basic.py:376-381: ``z = rng.normal((n, d))``; normalise rows;
    scale by ``rng.uniform(n) ** (1/d)``; un-whiten with B and c.
    """
    n_dim = z.shape[1]
    points = z / np.sqrt(np.sum(z**2, axis=1))[:, np.newaxis]
    points = points * (u[:, np.newaxis]**(1.0 / n_dim))
    return ell_transform(ell, points, inverse=True)


def ell_log_v(ell):
    """``Ellipsoid.log_v`` (bounds/basic.py:393-394)."""
    n_dim = len(ell['c'])
    return (np.linalg.slogdet(ell['B'])[1] + n_dim * np.log(2.) +
            n_dim * gammaln(1.5) - gammaln(n_dim / 2.0 + 1))


# --------------------------------------------------------------------------
# UnitCubeEllipsoidMixture  (bounds/basic.py:452-726)
# --------------------------------------------------------------------------

def mix_contains(mix, points):
    """``UnitCubeEllipsoidMixture.contains`` (bounds/basic.py:610-617)."""
    dim_cube = np.asarray(mix['dim_cube'], dtype=bool)
    in_bound = np.ones(points.shape[:-1], dtype=bool)
    if np.any(dim_cube):
        in_bound = in_bound & cube_contains(points[..., dim_cube])
    if mix['ell'] is not None:
        in_bound = in_bound & ell_contains(mix['ell'], points[..., ~dim_cube])
    return in_bound


def mix_transform(mix, points):
    """``UnitCubeEllipsoidMixture.transform`` (bounds/basic.py:584-592)."""
    dim_cube = np.asarray(mix['dim_cube'], dtype=bool)
    points_t = np.copy(points)
    if np.any(dim_cube):
        points_t[:, dim_cube] = points[:, dim_cube] * 2 - 1
    if mix['ell'] is not None:
        points_t[:, ~dim_cube] = ell_transform(
            mix['ell'], points[:, ~dim_cube])
    return points_t


def mix_sample_from(mix, cube_u, z, u):
    """``UnitCubeEllipsoidMixture.sample`` (bounds/basic.py:633-640).

    ``cube_u`` f64[n, n_cube] fills the cube dimensions, (z, u) feed the
    ellipsoid over the remaining ones; the reference draws the cube block
    first, then the ellipsoid's normals and uniforms.
    """
    dim_cube = np.asarray(mix['dim_cube'], dtype=bool)
    n = len(u) if u is not None else len(cube_u)
    points = np.zeros((n, len(dim_cube)))
    if np.any(dim_cube):
        points[:, dim_cube] = cube_u
    if mix['ell'] is not None:
        points[:, ~dim_cube] = ell_sample_from(mix['ell'], z, u)
    return points


def mix_log_v(mix):
    """``UnitCubeEllipsoidMixture.log_v`` (bounds/basic.py:651-655)."""
    return 0 if mix['ell'] is None else ell_log_v(mix['ell'])


# --------------------------------------------------------------------------
# Union  (bounds/union.py:43-450)
# --------------------------------------------------------------------------

def union_count(spec, points):
    """Overlap count ``n_bound = sum_k contains_k`` (bounds/union.py:316-317).

    Integer, must be bit-exact.
    """
    return np.sum([mix_contains(m, points) for m in spec['mixtures']],
                  axis=0).astype(np.int64)


def union_contains(spec, points):
    """``Union.contains`` (bounds/union.py:285-289)."""
    in_bound = np.any([mix_contains(m, points) for m in spec['mixtures']],
                      axis=0)
    if spec['unit']:
        in_bound = in_bound & cube_contains(points)
    return in_bound


def union_probabilities(spec):
    """Volume-proportional choice weights (bounds/union.py:308)."""
    log_v_all = np.asarray(spec['log_v_all'], dtype=float)
    return np.exp(log_v_all - logsumexp(log_v_all))


def union_accept(spec, points, r):
    """One ``Union.sample`` iteration after the draw (bounds/union.py:313-323).

    Given raw draws ``points`` (already in the reference's post-shuffle order
    where that matters) and the acceptance uniforms ``r`` (one per point that
    survives the cube filter, in order), return
    ``(in_cube, n_bound, accept)`` where ``n_bound``/``accept`` are defined on
    the in-cube points.  ``n_bound == 0`` gives ``p = -inf`` so the point is
    always accepted, exactly as in the reference.
    """
    if spec['unit']:
        in_cube = cube_contains(points)
    else:
        in_cube = np.ones(len(points), dtype=bool)
    kept = points[in_cube]
    n_bound = union_count(spec, kept)
    with np.errstate(divide='ignore'):
        p = 1 - 1.0 / n_bound
    accept = r[:len(kept)] > p
    return in_cube, n_bound, accept


def union_log_v(spec, n_sample, n_reject):
    """``Union.log_v`` (bounds/union.py:342-343)."""
    return logsumexp(spec['log_v_all']) + np.log(1.0 - n_reject / n_sample)


# --------------------------------------------------------------------------
# NeuralNetworkEmulator  (neural.py:35-187) -> sklearn MLPRegressor forward
# --------------------------------------------------------------------------

def mlp_forward(x, coefs, intercepts):
    """Forward pass of one ``MLPRegressor``.

    scikit-learn 1.9.0 ``neural_network/_multilayer_perceptron.py:189-224``
    (``_forward_pass_fast``): ``a = a @ W_i + b_i``; ReLU on all but the last
    layer; identity output; ``predict`` ravels the single output column
    (``:1751-1756``).
    """
    a = x
    n = len(coefs)
    for i in range(n):
        a = a @ coefs[i]
        a = a + intercepts[i]
        if i != n - 1:
            a = np.maximum(a, 0)
    return a.ravel() if a.shape[1] == 1 else a


def emulator_predict(emu, x):
    """``NeuralNetworkEmulator.predict`` (neural.py:114-116)."""
    xs = (x - emu['mean']) / emu['scale']
    return np.mean([mlp_forward(xs, w, b) for w, b in
                    zip(emu['coefs'], emu['intercepts'])], axis=0)


# --------------------------------------------------------------------------
# NeuralBound / NautilusBound  (bounds/neural.py, bounds/nautilus.py)
# --------------------------------------------------------------------------

def neural_contains(nb, points, return_score=False):
    """``NeuralBound.contains`` (bounds/neural.py:115-126)."""
    points = np.atleast_2d(points)
    in_bound = ell_contains(nb['ell'], points)
    score = np.full(len(points), np.nan)
    if np.any(in_bound) and nb['emulator'] is not None:
        points_t = ell_transform(nb['ell'], points)
        score[in_bound] = emulator_predict(nb['emulator'], points_t[in_bound])
        in_bound[in_bound] = (score[in_bound] >
                              nb['score_predict_min'] - 1e-9)
    if return_score:
        return in_bound, score
    return in_bound


def bound_contains(spec, points):
    """``NautilusBound.contains`` (bounds/nautilus.py:162-169) or the cube."""
    if spec['kind'] == 'cube':
        return cube_contains(points)
    in_bound = union_contains(spec, points)
    if len(spec['neural']) > 0:
        in_bound = in_bound & np.any(
            [neural_contains(nb, points) for nb in spec['neural']], axis=0)
    return in_bound


def neural_filter(spec, points):
    """The filter inside ``NautilusBound.sample`` (bounds/nautilus.py:217-218).
    """
    return np.any([neural_contains(nb, points) for nb in spec['neural']],
                  axis=0)


def bound_log_v(spec, n_sample_u, n_reject_u, n_sample_b, n_reject_b):
    """``NautilusBound.log_v`` (bounds/nautilus.py:260-261)."""
    if spec['kind'] == 'cube':
        return 0.0
    return (union_log_v(spec, n_sample_u, n_reject_u) +
            np.log(1.0 - n_reject_b / n_sample_b))


# --------------------------------------------------------------------------
# Shell bookkeeping  (sampler.py:650-730, 910-943)
# --------------------------------------------------------------------------

def shell_info(log_l, bound_log_v, shell_n_sample):
    """``Sampler.update_shell_info`` (sampler.py:925-943).

    Returns (shell_n, shell_log_v, shell_log_l, shell_n_eff).
    """
    shell_n = len(log_l)
    if shell_n > 0:
        shell_log_v = bound_log_v + np.log(shell_n / shell_n_sample)
        shell_log_l = logsumexp(log_l) - np.log(shell_n)
        if not np.all(log_l == -np.inf):
            shell_n_eff = np.exp(2 * logsumexp(log_l) - logsumexp(2 * log_l))
        else:
            shell_n_eff = len(log_l)
    else:
        shell_log_v, shell_log_l, shell_n_eff = -np.inf, np.nan, 0
    return shell_n, shell_log_v, shell_log_l, shell_n_eff


def log_z(shell_n, shell_log_l, shell_log_v):
    """``Sampler.log_z`` (sampler.py:690-694)."""
    if np.sum(shell_n) == 0:
        return None
    select = ~np.isnan(shell_log_l)
    return logsumexp(shell_log_l[select] + shell_log_v[select])


def n_eff(shell_log_l, shell_log_v, shell_n_eff):
    """``Sampler.n_eff`` (sampler.py:659-665)."""
    if np.all(shell_n_eff == 0):
        return 0
    select = shell_n_eff > 0
    sum_w = np.exp(shell_log_l + shell_log_v -
                   np.nanmax(shell_log_l + shell_log_v))[select]
    sum_w_sq = sum_w**2 / shell_n_eff[select]
    return np.sum(sum_w)**2 / np.sum(sum_w_sq)


def eta(shell_n, shell_log_l, shell_log_v, shell_n_eff):
    """``Sampler.eta`` (sampler.py:723-730)."""
    shell_log_z = shell_log_l + shell_log_v
    shell_eta = shell_n_eff / shell_n
    select = ~np.isnan(shell_log_l)
    shell_log_z = shell_log_z[select]
    shell_eta = shell_eta[select]
    return np.exp(2 * logsumexp(shell_log_z) - 2 * logsumexp(
        shell_log_z - 0.5 * np.log(shell_eta)))


def posterior_log_w(shell_n, shell_log_v, log_l_per_shell):
    """Importance weights of ``Sampler.posterior`` (sampler.py:601-606, 642).
    """
    log_v = np.repeat(shell_log_v - np.log(np.maximum(shell_n, 1)), shell_n)
    log_l = np.concatenate(log_l_per_shell)
    log_w = log_v + log_l
    return log_w - logsumexp(log_w)


# --------------------------------------------------------------------------
# Synthetic likelihoods -- NumPy face
# --------------------------------------------------------------------------

def loglike_gaussian(x, mu, inv_sigma2):
    """Isotropic Gaussian, log L = -0.5 * sum(((x - mu))^2) / sigma^2 + norm.

    Normalised so that the evidence over an unbounded prior is 1 (log Z = 0).
    """
    d = x.shape[-1]
    norm = 0.5 * d * np.log(inv_sigma2 / (2 * np.pi))
    return -0.5 * inv_sigma2 * np.sum((x - mu)**2, axis=-1) + norm


# --------------------------------------------------------------------------
# One batch of the full cycle, CPU (used as checker and as CPU baseline)
# --------------------------------------------------------------------------

def categorical_cdf(spec):
    """CDF used by the kernels for the per-proposal ellipsoid choice.

    The reference draws ``multinomial(1000, p)`` (bounds/union.py:308-309);
    N iid categorical draws with the same ``p`` have the same law.
    """
    cdf = np.cumsum(union_probabilities(spec))
    cdf[-1] = 1.0
    return cdf


def replay_integer_stream(n, offset, stream, seed, spec):
    """Replay block 0 of each proposal's Philox stream (exact in fp64).

    Returns (k_choice, r_accept, u_radial) for proposal indices
    ``offset .. offset+n-1``; mirrors ``nb200_rng.cuh``.
    """
    idx = np.arange(n, dtype=np.uint64) + np.uint64(offset)
    w0, w1, w2, w3 = philox.philox_block(idx, 0, stream, seed)
    if spec['kind'] == 'cube':
        k = np.zeros(n, dtype=np.int64)
    else:
        cdf = categorical_cdf(spec)
        k = np.minimum(np.searchsorted(cdf, philox.u01_32(w0), side='right'),
                       len(cdf) - 1)
    return k, philox.u01_32(w1), philox.u01_53(w2, w3)


def classify(spec, later, points, r, loglike=None):
    """Disposition of raw proposals, exactly as the reference would decide.

    points f64[N, d] are the raw union draws (any source), r f64[N] the
    acceptance uniforms.  Codes (shared with ``include/nautilus_b200.h``):
      0 rejected by the unit cube      (bounds/union.py:313-314)
      1 rejected by the overlap rule   (bounds/union.py:316-319)
      2 rejected by the neural filter  (bounds/nautilus.py:217-219)
      3 excluded by a later bound      (sampler.py:796-801)
      4 in shell (likelihood evaluated)
    Returns (code u8[N], n_bound i32[N], log_l f64[N] (nan unless code 4)).
    """
    n = len(points)
    code = np.zeros(n, dtype=np.uint8)
    nb_out = np.zeros(n, dtype=np.int32)
    log_l = np.full(n, np.nan)
    if spec['kind'] == 'cube':
        alive = np.ones(n, dtype=bool)
    else:
        in_cube = cube_contains(points) if spec['unit'] else np.ones(n, bool)
        idx = np.flatnonzero(in_cube)
        n_bound = union_count(spec, points[idx])
        nb_out[idx] = n_bound
        with np.errstate(divide='ignore'):
            accept = r[idx] > 1 - 1.0 / n_bound
        code[idx[~accept]] = 1
        idx = idx[accept]
        if len(spec['neural']) > 0 and len(idx) > 0:
            nn = neural_filter(spec, points[idx])
        else:
            nn = np.ones(len(idx), dtype=bool)
        code[idx[~nn]] = 2
        alive = np.zeros(n, dtype=bool)
        alive[idx[nn]] = True
    idx = np.flatnonzero(alive)
    excluded = np.zeros(len(idx), dtype=bool)
    for b in later:
        if len(idx) > 0:
            excluded |= bound_contains(b, points[idx])
    code[idx[excluded]] = 3
    idx = idx[~excluded]
    code[idx] = 4
    if loglike is not None and len(idx) > 0:
        log_l[idx] = loglike(points[idx])
    return code, nb_out, log_l


def lse_triple(log_l):
    """(max, sum exp(l-m), sum exp(2(l-m))) -- what ``shell_lse`` returns."""
    if len(log_l) == 0:
        return -np.inf, 0.0, 0.0
    m = np.max(log_l)
    if not np.isfinite(m):
        return m, 0.0, 0.0
    e = np.exp(log_l - m)
    return m, np.sum(e), np.sum(e * e)


# --------------------------------------------------------------------------
# CPU sampling path (reference algorithm end to end, explicit RNG)
# --------------------------------------------------------------------------

def union_sample_iteration(spec, rng, n_sample=1000):
    """One iteration of the loop in ``Union.sample`` (bounds/union.py:305-323)
    consuming ``rng`` in the reference's order.  Returns (accepted points,
    n_sample, n_reject_increment)."""
    p = union_probabilities(spec)
    n_per_bound = rng.multinomial(n_sample, p)
    chunks = []
    for mix, n in zip(spec['mixtures'], n_per_bound):
        dim_cube = np.asarray(mix['dim_cube'], dtype=bool)
        # Draw order of the reference: cube block first (basic.py:636), then
        # the ellipsoid's normals and radial uniforms (basic.py:376-379).
        cube_u = z = u = None
        if np.any(dim_cube):
            cube_u = rng.random(size=(n, int(np.sum(dim_cube))))
        if mix['ell'] is not None:
            z = rng.normal(size=(n, int(np.sum(~dim_cube))))
            u = rng.uniform(size=n)
        chunks.append(mix_sample_from(mix, cube_u, z, u))
    points = np.vstack(chunks)
    if spec['unit']:
        points = points[cube_contains(points)]
    rng.shuffle(points)
    n_bound = union_count(spec, points)
    with np.errstate(divide='ignore'):
        p = 1 - 1.0 / n_bound
    points = points[rng.random(size=len(points)) > p]
    return points, n_sample, n_sample - len(points)


def nautilus_sample(spec, rng, n_points, state=None):
    """``NautilusBound.sample`` without a pool (bounds/nautilus.py:213-222,
    239-241) on top of ``Union.sample`` (bounds/union.py:305-327), keeping
    both FIFO buffers and all four integer counters in ``state``."""
    d = spec['n_dim']
    if state is None:
        state = dict(u_points=np.zeros((0, d)), u_n_sample=0, u_n_reject=0,
                     points=np.zeros((0, d)), n_sample=0, n_reject=0)
    while len(state['points']) < n_points:
        while len(state['u_points']) < 1000:
            pts, ns, nr = union_sample_iteration(spec, rng)
            state['u_points'] = np.vstack([state['u_points'], pts])
            state['u_n_sample'] += ns
            state['u_n_reject'] += nr
        points = state['u_points'][:1000]
        state['u_points'] = state['u_points'][1000:]
        in_bound = neural_filter(spec, points)
        points = points[in_bound]
        state['points'] = np.vstack([state['points'], points])
        state['n_sample'] += 1000
        state['n_reject'] += 1000 - len(points)
    out = state['points'][:n_points]
    state['points'] = state['points'][n_points:]
    return out, state
