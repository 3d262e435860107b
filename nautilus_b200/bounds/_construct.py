"""Construction helpers: enclosing ellipsoids, 2-cluster mixtures, the
dimension search of the cube-ellipsoid mixture, ellipsoid overlap.

Bound *construction* happens between shells (SURVEY.md section 8, row a23 /
"next" f-1), on a few thousand live points.  Every piece has a device form
(the default) and a host NumPy form (``NB200_CONSTRUCT=host``; also the
cross-check of the tests).  The algorithms are written from their published
descriptions, not from the reference's code:

* minimum-volume enclosing ellipsoid: Khachiyan's first-order algorithm in
  the lifted space (Todd & Yildirim 2007) with Sherman-Morrison rank-one
  updates of the inverse and O(N d) updates of the Mahalanobis distances
  (device: the persistent cluster kernel behind ``ops.mvee_weights``); the
  reference (nautilus/bounds/basic.py:175-241) runs a batched variant with a
  full re-inversion per update.
* two-component Gaussian mixture by EM with k-means++ restarts (the reference
  calls scikit-learn's GaussianMixture(n_components=2, n_init=10),
  nautilus/bounds/union.py:185-187); device: all restarts in ONE launch,
  ``ops.gmm2_em`` (csrc/nb200_gmm.cu), or batched tensor operations for
  shapes outside that kernel's envelope.
* dimension search (nautilus/bounds/basic.py:497-512): the projected volume of
  every candidate dimension from one matrix product (``projection_scan``).
* ellipsoid overlap: Gilitschenski & Hanebeck 2012 (the test cited at
  nautilus/bounds/union.py:17); device: all pairs at once through a
  generalised eigen-decomposition (``overlap_k_min``); host: golden-section
  search.
"""

import hashlib
from collections import OrderedDict

import numpy as np

def construction_device():
    """Where the enclosing-ellipsoid iteration and the mixture EM run: the
    current CUDA device, or None (host NumPy) with NB200_CONSTRUCT=host."""
    import os
    if os.environ.get('NB200_CONSTRUCT', 'device') == 'host':
        return None
    from .._device import default_device
    return default_device()


_MVEE_CACHE = OrderedDict()      # the same live set is bounded several times
_MVEE_CACHE_SIZE = 32            # per new bound (nautilus.py:100-119)


def _khachiyan(points, max_updates, tol):
    """Weights u of Khachiyan's algorithm on well-conditioned points."""
    n, d = points.shape
    q = np.hstack([points, np.ones((n, 1))])
    u = np.full(n, 1.0 / n)
    v_inv = np.linalg.inv((q * u[:, None]).T @ q)
    g = np.einsum('ij,ij->i', q @ v_inv, q)
    for it in range(max_updates):
        j = int(np.argmax(g))
        g_max = g[j]
        if g_max <= (d + 1) * (1 + tol):
            break
        step = (g_max - (d + 1)) / ((d + 1) * (g_max - 1))
        beta = step / (1 - step)
        u *= 1 - step
        u[j] += step
        if it % 64 == 63:
            # refresh the inverse from scratch now and then (rounding drift)
            v_inv = np.linalg.inv((q * u[:, None]).T @ q)
            g = np.einsum('ij,ij->i', q @ v_inv, q)
            continue
        w = v_inv @ q[j]
        denom = 1 + beta * g_max
        # inverse and distances after V <- (1 - step) V + step q_j q_j^T
        v_inv = (v_inv - (beta / denom) * np.outer(w, w)) / (1 - step)
        g = (g - (beta / denom) * (q @ w)**2) / (1 - step)
    return u


def enclosing_ellipsoid(points, max_updates=1500, tol=1e-3, device=None):
    """Approximate MVEE of ``points`` [N, d].

    ``max_updates`` exact rank-one updates (the distances of ALL points are
    brought up to date after every update).  The reference stops after at
    most 100 x 20 = 2000 updates whose candidates come from distances that
    are up to 20 updates stale (nautilus/bounds/basic.py:214-228); 1500 exact
    updates still give the tighter ellipsoid (d = 30 golden: log V 0.05 below
    the reference's; 1000 updates would match it), and the iteration never
    reaches ``tol`` at these sizes anyway -- its cost is the update count.

    With ``device`` (a CUDA device) the Khachiyan iteration runs in the
    persistent kernel ``k_mvee`` (csrc/nb200_construct.cu) and the O(N d^2)
    pre- and post-processing as fp64 tensor operations on that device;
    otherwise everything is NumPy on the host.

    Returns ``(c, A, A_inv)`` with ``max_i (x_i - c)^T A (x_i - c) == 1``
    (the convention of nautilus/bounds/basic.py:233-241).  The MVEE is affine
    equivariant, so the iteration runs on points whitened by their sample
    covariance, which keeps it stable for extremely elongated clouds.
    """
    points = np.ascontiguousarray(points, dtype=float)
    n, d = points.shape
    key = (points.shape, max_updates, tol,
           hashlib.blake2b(points.tobytes(), digest_size=16).digest())
    if key in _MVEE_CACHE:
        _MVEE_CACHE.move_to_end(key)
        c, a, a_inv = _MVEE_CACHE[key]
        return c.copy(), a.copy(), a_inv.copy()
    if device is not None:
        result = _enclosing_ellipsoid_device(points, max_updates, tol, device)
        _MVEE_CACHE[key] = result
        if len(_MVEE_CACHE) > _MVEE_CACHE_SIZE:
            _MVEE_CACHE.popitem(last=False)
        return tuple(r.copy() for r in result)
    mu = np.mean(points, axis=0)
    cov = np.atleast_2d(np.cov(points, rowvar=False))
    cov = cov + np.eye(d) * 1e-14 * max(np.trace(cov) / d, 1e-300)
    chol = np.linalg.cholesky(cov)
    white = np.linalg.solve(chol, (points - mu).T).T
    u = _khachiyan(white, max_updates, tol)
    c = u @ points
    diff = points - c
    a_inv = (diff * u[:, None]).T @ diff
    a_inv = np.atleast_2d(0.5 * (a_inv + a_inv.T))
    a = np.linalg.inv(a_inv)
    a = 0.5 * (a + a.T)
    scale = np.max(np.einsum('ij,jk,ik->i', diff, a, diff))
    result = (np.atleast_1d(c), a / scale, a_inv * scale)
    _MVEE_CACHE[key] = result
    if len(_MVEE_CACHE) > _MVEE_CACHE_SIZE:
        _MVEE_CACHE.popitem(last=False)
    return tuple(r.copy() for r in result)


def _enclosing_ellipsoid_device(points, max_updates, tol, device):
    """``enclosing_ellipsoid`` with the points on ``device``: same steps, the
    sequential loop in one persistent CUDA kernel."""
    import torch
    from .. import ops
    x = torch.from_numpy(points).to(device)
    n, d = x.shape
    mu = x.mean(dim=0)
    xc = x - mu
    cov = xc.T @ xc / (n - 1)
    eye = torch.eye(d, dtype=torch.float64, device=x.device)
    cov = cov + eye * 1e-14 * max(float(torch.trace(cov)) / d, 1e-300)
    chol = torch.linalg.cholesky(cov)
    white_t = torch.linalg.solve_triangular(chol, xc.T, upper=False)
    u, _ = ops.mvee_weights(white_t.contiguous(), max_updates, tol)
    c = u @ x
    diff = x - c
    a_inv = (diff * u[:, None]).T @ diff
    a_inv = 0.5 * (a_inv + a_inv.T)
    a = torch.linalg.inv(a_inv)
    a = 0.5 * (a + a.T)
    scale = torch.max(((diff @ a) * diff).sum(dim=1))
    return (np.atleast_1d(c.cpu().numpy()), (a / scale).cpu().numpy(),
            (a_inv * scale).cpu().numpy())


# --------------------------------------------------------------------------

def _log_gauss(x, mean, cov):
    d = x.shape[1]
    chol = np.linalg.cholesky(cov)
    sol = np.linalg.solve(chol, (x - mean).T)
    return (-0.5 * np.sum(sol**2, axis=0) - np.sum(np.log(np.diag(chol))) -
            0.5 * d * np.log(2 * np.pi))


def two_gaussians(x, rng, n_init=10, max_iter=100, tol=1e-3, reg=1e-6):
    """Fit a 2-component full-covariance Gaussian mixture by EM.

    Returns the per-point log joint density ``log(w_k N(x | mu_k, C_k))`` as
    an array [N, 2] for the best of ``n_init`` restarts.
    """
    x = np.asarray(x, dtype=float)
    n, d = x.shape
    best, best_ll = None, -np.inf
    eye = np.eye(d) * reg
    for _ in range(n_init):
        # k-means++ seeding followed by a hard assignment
        first = x[rng.integers(n)]
        dist = np.sum((x - first)**2, axis=1)
        second = x[rng.choice(n, p=dist / np.sum(dist))]
        resp = (np.sum((x - second)**2, axis=1) < dist).astype(float)
        resp = np.stack([1 - resp, resp], axis=1)
        ll_old = -np.inf
        log_p = None
        for _ in range(max_iter):
            nk = resp.sum(axis=0) + 1e-12
            if np.any(nk < d + 1):
                break
            weights = nk / n
            log_p = np.empty((n, 2))
            try:
                for k in range(2):
                    mean = resp[:, k] @ x / nk[k]
                    diff = x - mean
                    cov = (diff * resp[:, k][:, None]).T @ diff / nk[k] + eye
                    log_p[:, k] = _log_gauss(x, mean, cov) + np.log(weights[k])
            except np.linalg.LinAlgError:
                log_p = None
                break
            m = np.max(log_p, axis=1, keepdims=True)
            norm = m + np.log(np.sum(np.exp(log_p - m), axis=1, keepdims=True))
            resp = np.exp(log_p - norm)
            ll = float(np.mean(norm))
            if abs(ll - ll_old) < tol:
                break
            ll_old = ll
        if log_p is not None and ll_old > best_ll:
            best, best_ll = log_p, ll_old
    if best is None:      # degenerate data: split along the widest direction
        axis = np.argmax(np.var(x, axis=0))
        side = x[:, axis] > np.median(x[:, axis])
        best = np.stack([np.where(side, -1.0, 0.0),
                         np.where(side, 0.0, -1.0)], axis=1)
    return best


def _median_split(x):
    """Degenerate data: split along the widest direction."""
    axis = np.argmax(np.var(x, axis=0))
    side = x[:, axis] > np.median(x[:, axis])
    return np.stack([np.where(side, -1.0, 0.0),
                     np.where(side, 0.0, -1.0)], axis=1)


def two_gaussians_batched(x, rng, n_init=10, max_iter=100, tol=1e-3,
                          reg=1e-6, device='cuda', fused=None):
    """``two_gaussians`` with the EM of all ``n_init`` restarts advanced
    together as batched fp64 tensor operations on ``device``.

    Same algorithm, same consumption of ``rng`` (the k-means++ seeding runs on
    the host), same rules for abandoning / ranking restarts; one EM iteration
    is ~25 launches whatever N, d and n_init are, instead of ~40 NumPy calls
    per restart (0.8 s -> ~15 ms for 2000 x 30-D, 30 s -> ~50 ms for
    10000 x 100-D).  Returns the [N, 2] log joint densities as NumPy.

    ``fused`` (default: on a CUDA device, when the shape fits): the whole EM
    of all restarts is ONE kernel launch (``ops.gmm2_em``, a thread-block
    cluster per restart) instead of ~25 launches per iteration.
    """
    import torch
    x_h = np.ascontiguousarray(x, dtype=float)
    n, d = x_h.shape
    resp0 = np.empty((n_init, 2, n))
    for r in range(n_init):
        first = x_h[rng.integers(n)]
        dist = np.sum((x_h - first)**2, axis=1)
        second = x_h[rng.choice(n, p=dist / np.sum(dist))]
        near = (np.sum((x_h - second)**2, axis=1) < dist).astype(float)
        resp0[r, 0], resp0[r, 1] = 1 - near, near
    dev = torch.device(device)
    xt = torch.from_numpy(x_h).to(dev)
    if fused is None:
        fused = dev.type == 'cuda'
    if fused:
        from .. import ops
        fused = ops.gmm2_applicable(n, d)
    if fused:
        # one launch: a thread-block cluster per restart runs its EM start to
        # finish (csrc/nb200_gmm.cu); the host reads R scores and the winner
        labels = torch.from_numpy(
            np.ascontiguousarray(resp0[:, 1, :]).astype(np.uint8)).to(dev)
        log_p, score, _ = ops.gmm2_em(xt.contiguous(), labels,
                                      max_iter=max_iter, tol=tol, reg=reg)
        score_h = score.cpu().numpy()
        best, best_ll = None, -np.inf
        for r in range(n_init):
            if score_h[r] > best_ll:
                best, best_ll = r, score_h[r]
        if best is None:
            return _median_split(x_h)
        return log_p[best].transpose(0, 1).contiguous().cpu().numpy()
    resp = torch.from_numpy(resp0).to(dev)                      # [R, 2, N]
    eye = torch.eye(d, dtype=torch.float64, device=dev) * reg
    done = torch.zeros(n_init, dtype=torch.bool, device=dev)
    valid = torch.zeros(n_init, dtype=torch.bool, device=dev)
    ll_old = torch.full((n_init,), -np.inf, dtype=torch.float64, device=dev)
    keep = torch.zeros((n_init, 2, n), dtype=torch.float64, device=dev)
    const = 0.5 * d * np.log(2 * np.pi)
    for it in range(max_iter):
        nk = resp.sum(dim=2) + 1e-12                             # [R, 2]
        starved = (nk < d + 1).any(dim=1)
        mean = torch.matmul(resp, xt) / nk[:, :, None]           # [R, 2, d]
        diff = xt[None, None] - mean[:, :, None, :]              # [R, 2, N, d]
        cov = torch.matmul((diff * resp[..., None]).transpose(-1, -2),
                           diff) / nk[:, :, None, None] + eye
        # starved components may have singular covariances; they are
        # abandoned anyway, keep the factorisation finite
        cov = torch.where(starved[:, None, None, None], eye / reg, cov)
        chol, info = torch.linalg.cholesky_ex(cov)
        failed = (info != 0).any(dim=1)
        chol = torch.where(failed[:, None, None, None], eye / reg, chol)
        sol = torch.linalg.solve_triangular(chol, diff.transpose(-1, -2),
                                            upper=False)         # [R, 2, d, N]
        log_p = (-0.5 * (sol * sol).sum(dim=2) -
                 torch.log(torch.diagonal(chol, dim1=-2, dim2=-1)).sum(
                     dim=-1)[..., None] - const +
                 torch.log(nk / n)[..., None])                   # [R, 2, N]
        m = log_p.max(dim=1, keepdim=True).values
        norm = m + torch.log(torch.exp(log_p - m).sum(dim=1, keepdim=True))
        ll = norm.mean(dim=2)[:, 0]
        step = ~done & ~starved & ~failed
        # a failed factorisation discards the restart; a starved component
        # stops it with what it had (two_gaussians: `break`)
        valid = torch.where(~done & failed & ~starved,
                            torch.zeros_like(valid), valid)
        valid = valid | step
        keep = torch.where(step[:, None, None], log_p, keep)
        resp = torch.where(step[:, None, None], torch.exp(log_p - norm), resp)
        conv = step & ((ll - ll_old).abs() < tol)
        ll_old = torch.where(step & ~conv, ll, ll_old)
        done = done | ~step | conv
        # (restarts that are done stay frozen, so looking every fourth
        # iteration changes nothing but the number of host synchronisations)
        if it % 4 == 3 and bool(done.all()):
            break
    score = torch.where(valid, ll_old, torch.full_like(ll_old, -np.inf))
    score_h = score.cpu().numpy()
    best, best_ll = None, -np.inf
    for r in range(n_init):              # first strict maximum, like the loop
        if score_h[r] > best_ll:
            best, best_ll = r, score_h[r]
    if best is None:
        axis = np.argmax(np.var(x_h, axis=0))
        side = x_h[:, axis] > np.median(x_h[:, axis])
        return np.stack([np.where(side, -1.0, 0.0),
                         np.where(side, 0.0, -1.0)], axis=1)
    return keep[best].transpose(0, 1).contiguous().cpu().numpy()


# --------------------------------------------------------------------------

def projection_scan(points, c, a):
    """Dimension search of ``UnitCubeEllipsoidMixture.compute``
    (nautilus/bounds/basic.py:497-512): for EVERY dimension i of the ellipsoid
    ``(x - c)^T A (x - c) <= 1``, the log volume factor of its projection
    along i, rescaled so that it still encloses the projected points.

    The reference inverts an (m-1) x (m-1) matrix and evaluates a quadratic
    form over all points for each candidate: O(m) inversions and O(n m^3)
    flops per step of the search.  With y = x - c, G = y A and Q = rowsum(G y)
    the Schur complement gives all candidates at once,

        y_k^T A_proj,i y_k = Q - G_i^2 / A_ii,
        log det (A^-1)_kk  = log A_ii - log det A,

    i.e. ONE [n, m] x [m, m] product: O(n m^2) per step.  ``points`` [n, m]
    and ``c``, ``a`` are tensors of one device (fp64); returns the tensor
    ``0.5 * log det(scale_i (A^-1)_kk)`` [m] -- what the reference calls
    ``log_v`` up to the constant 1/2, which does not move its argmin.
    """
    import torch
    y = points - c
    g = y @ a
    q = (g * y).sum(dim=1)
    diag = torch.diagonal(a)
    scale = (q[:, None] - g * g / diag).amax(dim=0)
    m = a.shape[0]
    return 0.5 * ((m - 1) * torch.log(scale) + torch.log(diag) -
                  torch.linalg.slogdet(a)[1])


def overlap_k_min(c, a_inv, pairs, levels=4, grid=1024):
    """min over s in (0, 1) of
        K(s) = 1 - d^T (A_1^-1 / (1 - s) + A_2^-1 / s)^-1 d,  d = c_1 - c_2,
    for all ``pairs`` of ellipsoids at once.  ``c`` [K, m] and ``a_inv``
    [K, m, m] are tensors of one device.

    One generalised eigen-decomposition per pair turns K into a scalar
    rational function: with A_1^-1 = L L^T, L^-1 A_2^-1 L^-T = U diag(lam) U^T
    and v = U^T L^-1 d,

        K(s) = 1 - sum_i v_i^2 s (1 - s) / (s + lam_i (1 - s)),

    which is evaluated on a grid of s for all pairs in one tensor expression
    and refined ``levels`` times around the minimum (K is convex in s): four
    levels of 1024 points bracket the minimiser to 1e-11, a few launches in
    total instead of a d x d solve per function evaluation.
    """
    import torch
    i1 = torch.as_tensor([p[0] for p in pairs], device=c.device)
    i2 = torch.as_tensor([p[1] for p in pairs], device=c.device)
    d = c[i1] - c[i2]                                          # [P, m]
    chol = torch.linalg.cholesky(a_inv[i1])
    w = torch.linalg.solve_triangular(chol, a_inv[i2], upper=False)
    w = torch.linalg.solve_triangular(chol, w.transpose(-1, -2), upper=False)
    lam, u = torch.linalg.eigh(0.5 * (w + w.transpose(-1, -2)))
    ld = torch.linalg.solve_triangular(chol, d[..., None], upper=False)
    v2 = (u.transpose(-1, -2) @ ld)[..., 0]**2                 # [P, m]
    lo = torch.full((len(pairs), ), 1e-9, dtype=c.dtype, device=c.device)
    hi = 1 - lo
    t = torch.linspace(0, 1, grid, dtype=c.dtype, device=c.device)
    k_min = None
    for _ in range(levels):
        s = lo[:, None] + (hi - lo)[:, None] * t               # [P, G]
        k = 1 - (v2[:, None, :] * (s * (1 - s))[..., None] /
                 (s[..., None] + lam[:, None, :] * (1 - s)[..., None])
                 ).sum(dim=-1)
        best = k.argmin(dim=1)
        k_min = k.gather(1, best[:, None])[:, 0]
        lo = s.gather(1, (best - 1).clamp(min=0)[:, None])[:, 0]
        hi = s.gather(1, (best + 1).clamp(max=grid - 1)[:, None])[:, 0]
    return k_min


def ellipsoids_overlap(ellipsoids, device=None):
    """True if any two ellipsoids intersect (nautilus/bounds/union.py:14-40).

    With ``device`` all pairs are tested at once on that device
    (``overlap_k_min``); the host form below is kept for
    ``NB200_CONSTRUCT=host`` and as the cross-check of the tests.
    """
    if device is not None and len(ellipsoids) > 1:
        import torch
        c = torch.as_tensor(np.stack([np.asarray(e.c) for e in ellipsoids]),
                            device=device)
        a = torch.as_tensor(np.stack([np.asarray(e.A) for e in ellipsoids]),
                            device=device)
        pairs = [(i, j) for i in range(len(ellipsoids))
                 for j in range(i + 1, len(ellipsoids))]
        k_min = overlap_k_min(c, torch.linalg.inv(a), pairs)
        return bool((k_min > 0).any().item())
    return _ellipsoids_overlap_host(ellipsoids)


def _ellipsoids_overlap_host(ellipsoids):
    """Golden-section form of the same test, NumPy.

    Two ellipsoids {(x-c_i)^T A_i (x-c_i) <= 1} are disjoint iff
    K(s) = 1 - d^T (A_1^{-1}/(1-s) + A_2^{-1}/s)^{-1} d  < 0 for some s in
    (0, 1), with d = c_1 - c_2; K is convex in s, so its minimum is found by
    golden-section search.
    """
    cs = [np.asarray(e.c) for e in ellipsoids]
    a_invs = [np.linalg.inv(e.A) for e in ellipsoids]
    inv_phi = (np.sqrt(5.0) - 1) / 2

    def k_of(s, d, m1, m2):
        return 1 - d @ np.linalg.solve(m1 / (1 - s) + m2 / s, d)

    for i in range(len(cs)):
        for j in range(i + 1, len(cs)):
            d = cs[i] - cs[j]
            lo, hi = 1e-9, 1 - 1e-9
            x1 = hi - inv_phi * (hi - lo)
            x2 = lo + inv_phi * (hi - lo)
            f1, f2 = k_of(x1, d, a_invs[i], a_invs[j]), k_of(
                x2, d, a_invs[i], a_invs[j])
            for _ in range(60):
                if f1 < f2:
                    hi, x2, f2 = x2, x1, f1
                    x1 = hi - inv_phi * (hi - lo)
                    f1 = k_of(x1, d, a_invs[i], a_invs[j])
                else:
                    lo, x1, f1 = x1, x2, f2
                    x2 = lo + inv_phi * (hi - lo)
                    f2 = k_of(x2, d, a_invs[i], a_invs[j])
            if min(f1, f2) > 0:
                return True
    return False
