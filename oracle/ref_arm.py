"""The UNMODIFIED reference (oracle/_ref/nautilus, placed there by
oracle/make_ref.sh) as the timed CPU arm.

TEST / BASELINE INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): imported by
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and by
``tests/test_oracle_golden.py``, never by ``nautilus_b200/``.

What is timed is the reference's own hot path, through its own public method:
``nautilus.Sampler.add_samples`` (nautilus/sampler.py:1093-1144) ->
``sample_shell`` (:751-830) -> ``NautilusBound.sample``
(bounds/nautilus.py:193-244) -> ``Union.sample`` (bounds/union.py:291-327) ->
``NeuralBound.contains`` (bounds/neural.py:99-126) -> scikit-learn
``MLPRegressor.predict`` -> likelihood (:832-908) -> ``update_shell_info``
(:910-943).  The bound is the config-2 bound the reference itself built
(tests/golden/cfg2_bound_d30.npz); it is put back into reference objects
attribute by attribute, the way the reference's own ``read`` methods do
(bounds/nautilus.py:329-380, nautilus/neural.py:153-187).
"""

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, '_ref')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'nautilus'))


def _import_reference():
    if not available():
        raise ImportError('oracle/_ref/nautilus is missing: run '
                          'oracle/make_ref.sh in the build container')
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import nautilus
    assert os.path.realpath(os.path.dirname(nautilus.__file__)).startswith(
        os.path.realpath(REF_ROOT)), 'imported another nautilus'
    return nautilus


def _ellipsoid(rb, ell, rng):
    obj = rb.Ellipsoid()
    obj.n_dim = len(ell['c'])
    obj.c = np.array(ell['c'])
    obj.B = np.array(ell['B'])
    obj.B_inv = np.array(ell['B_inv'])
    obj.A = obj.B_inv.T @ obj.B_inv
    obj.rng = rng
    return obj


def _mixture(rb, mix, rng):
    obj = rb.UnitCubeEllipsoidMixture()
    obj.dim_cube = np.array(mix['dim_cube'], dtype=bool)
    obj.n_dim = len(obj.dim_cube)
    n_cube = int(np.sum(obj.dim_cube))
    obj.cube = rb.UnitCube.compute(n_cube, rng=rng) if n_cube else None
    obj.ellipsoid = (None if mix['ell'] is None
                     else _ellipsoid(rb, mix['ell'], rng))
    obj.rng = rng
    return obj


def _emulator(emu):
    from nautilus.neural import NeuralNetworkEmulator
    from sklearn.neural_network import MLPRegressor
    obj = NeuralNetworkEmulator()
    obj.mean = np.array(emu['mean'])
    obj.scale = np.array(emu['scale'])
    obj.neural_networks = []
    for coefs, intercepts in zip(emu['coefs'], emu['intercepts']):
        net = MLPRegressor(hidden_layer_sizes=tuple(
            w.shape[1] for w in coefs[:-1]))
        net.coefs_ = [np.array(w) for w in coefs]
        net.intercepts_ = [np.array(b) for b in intercepts]
        net.n_layers_ = len(coefs) + 1
        net.n_outputs_ = 1
        net.n_features_in_ = coefs[0].shape[0]
        net.out_activation_ = 'identity'
        obj.neural_networks.append(net)
    return obj


def reference_bound(spec, rng):
    """A reference ``NautilusBound`` carrying the parameters of ``spec``."""
    _import_reference()
    from nautilus import bounds as rb
    from nautilus.bounds.neural import NeuralBound
    d = int(spec['n_dim'])
    union = rb.Union()
    union.n_dim = d
    union.enlarge_per_dim = 1.1
    union.n_points_min = d + 50
    union.cube = rb.UnitCube.compute(d, rng=rng) if spec['unit'] else None
    union.bounds = [_mixture(rb, m, rng) for m in spec['mixtures']]
    union.points_bounds = [np.zeros((0, d)) for _ in union.bounds]
    union.log_v_all = np.array(spec['log_v_all'], dtype=float)
    union.block = np.ones(len(union.bounds), dtype=bool)
    union.points = np.zeros((0, d))
    union.n_sample = 0
    union.n_reject = 0
    union.rng = rng
    bound = rb.NautilusBound()
    bound.n_dim = d
    bound.shift = None
    bound.neural_bounds = []
    for nb in spec['neural']:
        obj = NeuralBound()
        obj.n_dim = d
        obj.outer_bound = _ellipsoid(rb, nb['ell'], rng)
        obj.emulator = (None if nb['emulator'] is None
                        else _emulator(nb['emulator']))
        obj.score_predict_min = float(nb['score_predict_min'])
        bound.neural_bounds.append(obj)
    bound.outer_bound = union
    bound.rng = rng
    bound.points = np.zeros((0, d))
    bound.n_sample = 0
    bound.n_reject = 0
    return bound


def reference_sampler(spec, likelihood, log_l_min, seed=0, n_batch=1000,
                      pool=None):
    """A reference ``Sampler`` whose shell 1 is the bound of ``spec`` (shell 0
    is the unit cube its own ``add_bound`` creates), ready for
    ``add_samples(1)``."""
    nautilus = _import_reference()
    d = int(spec['n_dim'])
    sampler = nautilus.Sampler(lambda x: x, likelihood, n_dim=d, n_live=2000,
                               vectorized=True, pass_dict=False, seed=seed,
                               n_batch=n_batch, pool=pool)
    sampler.add_bound()                     # the unit cube (sampler.py:999)
    bound = reference_bound(spec, sampler.rng)
    # the bookkeeping of add_bound (sampler.py:1040-1057)
    sampler.bounds.append(bound)
    sampler.shell_n = np.append(sampler.shell_n, 0)
    sampler.shell_n_sample = np.append(sampler.shell_n_sample, 0)
    sampler.shell_n_eff = np.append(sampler.shell_n_eff, 0)
    sampler.shell_log_l = np.append(sampler.shell_log_l, np.nan)
    sampler.shell_log_v = np.append(sampler.shell_log_v, np.nan)
    sampler.shell_log_l_min = np.append(sampler.shell_log_l_min, log_l_min)
    sampler.points.append(np.zeros((0, d)))
    sampler.log_l.append(np.zeros(0))
    return sampler


def run_reference_cycles(spec, likelihood, log_l_min, n_raw, seed=0,
                         n_batch=1000, pool=None):
    """Call the reference's ``add_samples(1)`` until its outer union has
    consumed ``n_raw`` raw proposals.  Returns (raw proposals, seconds,
    dict of shell read-outs)."""
    sampler = reference_sampler(spec, likelihood, log_l_min, seed=seed,
                                n_batch=n_batch, pool=pool)
    union = sampler.bounds[1].outer_bound
    t0 = time.perf_counter()
    while union.n_sample < n_raw:
        sampler.add_samples(1)
    dt = time.perf_counter() - t0
    info = dict(shell_n=int(sampler.shell_n[1]),
                shell_log_l=float(sampler.shell_log_l[1]),
                shell_log_v=float(sampler.shell_log_v[1]),
                shell_n_eff=float(sampler.shell_n_eff[1]),
                n_like=int(sampler.n_like))
    return int(union.n_sample), dt, info
