"""Generate the golden fixtures in this directory from the REAL reference.

Run in the build container only (the reference is importable there):

    python tests/golden/make_golden.py

It imports johannesulf/nautilus v1.0.6 from ``/root/reference`` (read-only,
unmodified), builds bounds from seeded synthetic data with the reference's own
``compute`` class methods, calls the reference's ``contains / transform /
sample / predict / update_shell_info`` and stores inputs, exported parameters
and the reference's outputs as ``.npz`` files.  The fixtures pin both the CPU
oracle (``tests/test_oracle_golden.py``) and the CUDA kernels
(``tests/test_gpu_parity.py``).  Stochastic stages are made bit-reproducible
by cloning the generator state before the call and replaying the reference's
draw order.
"""

import os
import sys

import numpy as np

sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))

from nautilus import Sampler  # noqa: E402
from nautilus import bounds as rb  # noqa: E402
from nautilus.neural import NeuralNetworkEmulator  # noqa: E402

from nautilus_b200._pack import spec_to_flat  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- export reference objects as specs -----------------------------------

def ell_spec(ell):
    return dict(c=np.array(ell.c), B=np.array(ell.B), B_inv=np.array(ell.B_inv))


def mix_spec(mix, n_dim):
    if isinstance(mix, rb.Ellipsoid):
        return dict(dim_cube=np.zeros(n_dim, dtype=bool), ell=ell_spec(mix))
    return dict(dim_cube=np.array(mix.dim_cube),
                ell=None if mix.ellipsoid is None else ell_spec(mix.ellipsoid))


def emu_spec(emu):
    return dict(mean=np.array(emu.mean), scale=np.array(emu.scale),
                coefs=[[np.array(w) for w in net.coefs_]
                       for net in emu.neural_networks],
                intercepts=[[np.array(b) for b in net.intercepts_]
                            for net in emu.neural_networks])


def neural_spec(nb):
    return dict(ell=ell_spec(nb.outer_bound),
                emulator=None if nb.emulator is None else emu_spec(nb.emulator),
                score_predict_min=float(nb.score_predict_min))


def union_spec(union):
    return dict(kind='nautilus', n_dim=int(union.n_dim),
                unit=union.cube is not None,
                log_v_all=np.array(union.log_v_all, dtype=float),
                mixtures=[mix_spec(b, union.n_dim) for b in union.bounds],
                neural=[])


def nautilus_spec(bound):
    assert bound.shift is None
    spec = union_spec(bound.outer_bound)
    spec['neural'] = [neural_spec(nb) for nb in bound.neural_bounds]
    return spec


def clone_rng(rng):
    clone = np.random.default_rng()
    clone.bit_generator.state = rng.bit_generator.state
    return clone


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('{:28s} {:8.1f} KB'.format(name + '.npz',
                                     os.path.getsize(path) / 1024))


# ---- 1. single ellipsoids -------------------------------------------------

def make_ellipsoids():
    for d in (3, 10, 30):
        rng = np.random.default_rng(100 + d)
        # correlated, badly scaled cloud inside the unit cube
        L = np.tril(rng.normal(size=(d, d))) * 0.3 + np.eye(d)
        pts = 0.5 + 0.05 * rng.normal(size=(40 * d, d)) @ L.T / np.sqrt(d)
        ell = rb.Ellipsoid.compute(pts, enlarge_per_dim=1.1,
                                   rng=np.random.default_rng(0))
        # test points: half sampled inside (radii pile up at the boundary)
        clone = clone_rng(ell.rng)
        n = 1024 if d <= 3 else 384
        inside = ell.sample(n)
        z = clone.normal(size=(n, d))
        u = clone.uniform(size=n)
        # second half: the same directions pushed onto the surface +-0.1 %
        r_t = np.linalg.norm(ell.transform(inside), axis=1)[:, None]
        test = np.vstack([inside, ell.c + (inside - ell.c) / r_t * (
            1 + 1e-3 * rng.normal(size=(n, 1)))])
        save('ellipsoid_d{}'.format(d),
             c=ell.c, B=ell.B, B_inv=ell.B_inv, A=ell.A,
             z=z, u=u, sample=inside,
             points=test, transform=ell.transform(test),
             contains=ell.contains(test),
             inverse=ell.transform(ell.transform(test), inverse=True),
             log_v=ell.log_v)


# ---- 2. cube-ellipsoid mixture with cube dimensions ------------------------

def make_mixture():
    rng = np.random.default_rng(7)
    d, n = 6, 600
    pts = rng.random((n, d))
    pts[:, [1, 3, 4]] = 0.5 + 0.02 * rng.normal(size=(n, 3))
    mix = rb.UnitCubeEllipsoidMixture.compute(pts, rng=np.random.default_rng(0))
    assert np.any(mix.dim_cube) and not np.all(mix.dim_cube)
    clone = clone_rng(mix.ellipsoid.rng)
    m = 1024
    sample = mix.sample(m)
    cube_u = clone.random(size=(m, int(np.sum(mix.dim_cube))))
    z = clone.normal(size=(m, int(np.sum(~mix.dim_cube))))
    u = clone.uniform(size=m)
    test = np.vstack([sample, rng.random((m, d)) * 1.2 - 0.1])
    flat = spec_to_flat(dict(kind='nautilus', n_dim=d, unit=True,
                             log_v_all=np.array([mix.log_v]),
                             mixtures=[mix_spec(mix, d)], neural=[]))
    save('mixture_d6', cube_u=cube_u, z=z, u=u, sample=sample, points=test,
         contains=mix.contains(test), transform=mix.transform(test),
         log_v=mix.log_v, **flat)


# ---- 3. union of K overlapping ellipsoids, Union.sample replay -------------

def make_union():
    rng = np.random.default_rng(3)
    d = 5
    centres = np.array([[.3] * d, [.5] * d, [.62] * d])
    pts = np.vstack([c + 0.05 * rng.normal(size=(400, d)) for c in centres])
    union = rb.Union.compute(pts, enlarge_per_dim=1.1, n_points_min=d + 20,
                             bound_class=rb.UnitCubeEllipsoidMixture,
                             rng=np.random.default_rng(0))
    while len(union.bounds) < 4 and union.split():
        pass
    K = len(union.bounds)
    assert K >= 3, K
    spec = union_spec(union)
    # replay one Union.sample iteration (1000 raw draws)
    union.reset(np.random.default_rng(11))
    clone = clone_rng(union.rng)
    accepted = union.sample(1)      # exactly one loop iteration
    accepted = np.vstack([accepted, union.points])
    p = np.exp(np.array(union.log_v_all) -
               __import__('scipy.special').special.logsumexp(union.log_v_all))
    n_per_bound = clone.multinomial(1000, p)
    k_assign, zs, us, cus = [], [], [], []
    raw = []
    for k, (b, n) in enumerate(zip(union.bounds, n_per_bound)):
        nc = int(np.sum(b.dim_cube))
        cu = clone.random(size=(n, nc)) if nc > 0 else np.zeros((n, 0))
        if b.ellipsoid is not None:
            z = clone.normal(size=(n, d - nc))
            u = clone.uniform(size=n)
        else:
            z, u = np.zeros((n, 0)), np.zeros(n)
        k_assign.append(np.repeat(k, n))
        # pad to common width d so everything fits one array
        zs.append(np.pad(z, ((0, 0), (0, d - z.shape[1]))))
        cus.append(np.pad(cu, ((0, 0), (0, d - cu.shape[1]))))
        us.append(u)
    k_assign = np.concatenate(k_assign)
    zs, cus, us = np.vstack(zs), np.vstack(cus), np.concatenate(us)
    # raw points in draw order, from the reference's own sample methods fed
    # by a second clone
    clone2 = clone_rng(np.random.default_rng(11))
    clone2.multinomial(1000, p)
    for b, n in zip(union.bounds, n_per_bound):
        b.reset(clone2)
        raw.append(b.sample(n))
        b.reset(union.rng)
    raw = np.vstack(raw)
    in_cube = union.cube.contains(raw)
    perm = np.arange(int(np.sum(in_cube)))
    clone.shuffle(perm)
    r = clone.random(size=len(perm))
    shuffled = raw[in_cube][perm]
    n_bound = np.sum([b.contains(shuffled) for b in union.bounds], axis=0)
    acc = r > 1 - 1.0 / n_bound
    assert np.array_equal(shuffled[acc], accepted)
    assert union.n_sample == 1000
    assert union.n_reject == 1000 - len(accepted)
    test = np.vstack([raw, rng.random((1000, d))])
    save('union_d5', k_assign=k_assign, z=zs, cube_u=cus, u=us, raw=raw,
         in_cube=in_cube, perm=perm, r=r, n_bound=n_bound, accept=acc,
         accepted=accepted, n_sample=union.n_sample, n_reject=union.n_reject,
         points=test, contains=union.contains(test),
         count=np.sum([b.contains(test) for b in union.bounds], axis=0),
         log_v=union.log_v, **spec_to_flat(spec))


# ---- 4. NautilusBound on the reference test fixture (4-D) -------------------

def make_nautilus_4d():
    np.random.seed(0)
    points = np.random.random(size=(500, 4))       # tests/test_bounds.py:36-43
    log_l = -np.linalg.norm(points - 0.5, axis=1)
    log_l_min = np.median(log_l)
    bound = rb.NautilusBound.compute(
        points, log_l, log_l_min, np.log(0.5), n_networks=2,
        rng=np.random.default_rng(0))
    spec = nautilus_spec(bound)
    rng = np.random.default_rng(5)
    test = rng.random((4096, 4))
    nb = bound.neural_bounds[0]
    t = nb.outer_bound.transform(test)
    save('nautilus_d4', points=test, contains=bound.contains(test),
         union_contains=bound.outer_bound.contains(test),
         neural_contains=nb.contains(test),
         ell_contains=nb.outer_bound.contains(test),
         predict=nb.emulator.predict(t),
         train_points=points, train_log_l=log_l, log_l_min=log_l_min,
         **spec_to_flat(spec))


# ---- 5. config-2 bound: 30-D Gaussian, n_live=2000, 4 nets ------------------

def gauss30(x):
    return -0.5 * np.sum((x - 0.5)**2, axis=-1) / 0.1**2


def make_cfg2():
    rng = np.random.default_rng(0)
    d, m, n_live = 30, 6000, 2000
    z = rng.normal(size=(m, d))
    z /= np.linalg.norm(z, axis=1)[:, None]
    pts = 0.5 + 0.4 * z * rng.random((m, 1))**(1.0 / d)
    log_l = gauss30(pts)
    log_l_min = np.sort(log_l)[-n_live]
    log_v_target = np.log(n_live / m) + (
        d * np.log(0.4) + 0.5 * d * np.log(np.pi) -
        __import__('scipy.special').special.gammaln(d / 2 + 1))
    bound = rb.NautilusBound.compute(
        pts, log_l, log_l_min, log_v_target, n_networks=4,
        rng=np.random.default_rng(1))
    spec = nautilus_spec(bound)
    bound.sample(2000, return_points=False)
    test = np.vstack([bound.points[:768],
                      0.5 + 0.45 * (rng.random((768, d)) - 0.5)])
    nb = bound.neural_bounds[0]
    t = nb.outer_bound.transform(test)
    save('cfg2_bound_d30', points=test, contains=bound.contains(test),
         union_contains=bound.outer_bound.contains(test),
         neural_contains=nb.contains(test),
         ell_contains=nb.outer_bound.contains(test),
         predict=nb.emulator.predict(t),
         log_l_min=log_l_min, log_v_target=log_v_target,
         ref_log_v=bound.log_v, ref_n_sample=bound.n_sample,
         ref_n_reject=bound.n_reject,
         ref_u_n_sample=bound.outer_bound.n_sample,
         ref_u_n_reject=bound.outer_bound.n_reject,
         **spec_to_flat(spec))


# ---- 5b. the config-5 bound: 100-D correlated Gaussian --------------------

def make_cfg5():
    """d = 100, N(0.5, sigma^2 [(1-rho) I + rho 11^T]),
    sigma = 0.05, rho = 0.5, n_live = 10 000; the fixed bound the reference
    builds from 40 000 seeded points drawn from the target (minutes of CPU:
    `python tests/golden/make_golden.py cfg5`)."""
    from scipy.special import gammaln
    from scipy.stats import chi2
    from nautilus_b200 import likelihoods
    d, m, n_live = 100, 40000, 10000
    like = likelihoods.EquicorrelatedGaussian(d)
    rng = np.random.default_rng(0)
    sigma, rho = like.sigma, like.rho
    # x = mu + sigma (sqrt(1-rho) e + sqrt(rho) g 1), e ~ N(0, I), g ~ N(0, 1)
    pts = 0.5 + sigma * (np.sqrt(1 - rho) * rng.normal(size=(m, d)) +
                         np.sqrt(rho) * rng.normal(size=(m, 1)))
    assert np.all((pts > 0) & (pts < 1))
    log_l = like(pts)
    log_l_min = np.sort(log_l)[-n_live]
    # volume of the likelihood contour holding n_live / m of the mass
    r2 = chi2.ppf(n_live / m, d)
    log_det = (d * np.log(sigma**2) + (d - 1) * np.log(1 - rho) +
               np.log(1 + (d - 1) * rho))
    log_v_target = (0.5 * log_det + 0.5 * d * np.log(r2) +
                    0.5 * d * np.log(np.pi) - gammaln(d / 2 + 1))
    bound = rb.NautilusBound.compute(
        pts, log_l, log_l_min, log_v_target, n_networks=4,
        rng=np.random.default_rng(1))
    spec = nautilus_spec(bound)
    bound.sample(2000, return_points=False)
    test = np.vstack([bound.points[:512], pts[:512]])
    nb = bound.neural_bounds[0]
    t = nb.outer_bound.transform(test)
    save('cfg5_bound_d100', points=test, contains=bound.contains(test),
         union_contains=bound.outer_bound.contains(test),
         neural_contains=nb.contains(test),
         ell_contains=nb.outer_bound.contains(test),
         predict=nb.emulator.predict(t),
         log_l_min=log_l_min, log_v_target=log_v_target,
         ref_log_v=bound.log_v, ref_n_sample=bound.n_sample,
         ref_n_reject=bound.n_reject,
         ref_u_n_sample=bound.outer_bound.n_sample,
         ref_u_n_reject=bound.outer_bound.n_reject,
         **spec_to_flat(spec))


# ---- 6. shell bookkeeping from a real (small) reference run -----------------

def make_shells():
    def likelihood(x):
        return -0.5 * np.sum((x - 0.5)**2, axis=-1) / 0.1**2

    sampler = Sampler(lambda x: x, likelihood, n_dim=2, n_live=300,
                      vectorized=True, seed=0)
    sampler.run(n_eff=2000, verbose=False)
    arrays = dict(n_shells=len(sampler.bounds),
                  shell_n=sampler.shell_n,
                  shell_n_sample=sampler.shell_n_sample,
                  shell_n_eff=sampler.shell_n_eff,
                  shell_log_l=sampler.shell_log_l,
                  shell_log_v=sampler.shell_log_v,
                  shell_log_l_min=sampler.shell_log_l_min,
                  bound_log_v=np.array([b.log_v for b in sampler.bounds],
                                       dtype=float),
                  log_z=sampler.log_z, n_eff=sampler.n_eff, eta=sampler.eta)
    for i, ll in enumerate(sampler.log_l):
        arrays['log_l_{}'.format(i)] = ll
    pts, log_w, log_l = sampler.posterior()
    arrays['posterior_log_w'] = log_w
    # edge cases of update_shell_info (sampler.py:935-943)
    arrays['edge_all_minus_inf'] = np.full(5, -np.inf)
    save('shells_d2', **arrays)


# ---- 7. emulator on the reference's own test problem ------------------------

def make_emulator():
    np.random.seed(0)                               # tests/test_neural.py:9-15
    n_dim, n_points = 5, 1000
    x = np.random.random((n_points, n_dim))
    y = np.linalg.norm(x - 0.5, axis=1)
    y = np.argsort(np.argsort(y)) / float(len(y))
    emu = NeuralNetworkEmulator.train(x, y, n_networks=2, pool=None)
    flat = {}
    es = emu_spec(emu)
    flat['mean'], flat['scale'] = es['mean'], es['scale']
    for n in range(2):
        for i in range(4):
            flat['W{}_{}'.format(n, i)] = es['coefs'][n][i]
            flat['b{}_{}'.format(n, i)] = es['intercepts'][n][i]
    flat['n_iter'] = np.array([net.n_iter_ for net in emu.neural_networks])
    flat['loss'] = np.array([net.loss_ for net in emu.neural_networks])
    save('emulator_d5', x=x, y=y, predict=emu.predict(x), **flat)


def make_fit_stats():
    """scikit-learn's fit on the emulator_d5 data, 10 seeds (the reference's
    train_network, nautilus/neural.py:10-32, with random_state = 0..9): final
    training loss, epochs, and the rmse of the prediction against the target
    -- what tests/test_gpu_bounds_api.py compares the CUDA trainer with.
    40 seeds: the spread between seeds is a factor of 8 in the loss, ten
    would not pin a median."""
    from nautilus.neural import train_network
    with np.load(os.path.join(HERE, 'emulator_d5.npz')) as f:
        x, y = f['x'], f['y']
    xs = (x - np.mean(x, axis=0)) / np.std(x, axis=0)
    kwargs = dict(hidden_layer_sizes=(100, 50, 20), alpha=0,
                  learning_rate_init=1e-2, max_iter=10000, tol=0,
                  n_iter_no_change=10)
    loss, n_iter, rmse = [], [], []
    for seed in range(40):
        net = train_network(xs, y, kwargs, seed)
        loss.append(net.loss_)
        n_iter.append(net.n_iter_)
        rmse.append(np.sqrt(np.mean((net.predict(xs) - y)**2)))
    save('emulator_d5_sklearn_fits', loss=np.array(loss),
         n_iter=np.array(n_iter), rmse=np.array(rmse))
    print(loss, n_iter, rmse)


if __name__ == '__main__':
    if sys.argv[1:] == ['fits']:
        make_fit_stats()
        sys.exit(0)
    if sys.argv[1:] == ['cfg5']:
        make_cfg5()
        sys.exit(0)
    make_ellipsoids()
    make_mixture()
    make_union()
    make_nautilus_4d()
    make_cfg2()
    make_shells()
    make_emulator()
