"""Pin the CPU oracle against outputs of the real reference (tests/golden).

The fixtures were produced by ``tests/golden/make_golden.py`` which imports
johannesulf/nautilus v1.0.6; here the oracle has to reproduce them without the
reference being present.  Deterministic stages must match bit-for-bit (the
oracle issues the same NumPy calls as the reference).
"""

import numpy as np
import pytest
from scipy.special import gamma

from nautilus_b200._pack import flat_to_spec
from oracle import nautilus_oracle as orc
from oracle import philox


@pytest.mark.parametrize('d', [3, 10, 30])
def test_ellipsoid_matches_reference(golden, d):
    g = golden('ellipsoid_d{}'.format(d))
    ell = dict(c=g['c'], B=g['B'], B_inv=g['B_inv'])
    assert np.array_equal(orc.ell_transform(ell, g['points']), g['transform'])
    assert np.array_equal(orc.ell_contains(ell, g['points']), g['contains'])
    assert np.array_equal(
        orc.ell_transform(ell, g['transform'], inverse=True), g['inverse'])
    assert np.array_equal(orc.ell_sample_from(ell, g['z'], g['u']),
                          g['sample'])
    assert orc.ell_log_v(ell) == g['log_v']
    # about half of the probes sit just outside: the fixture is not trivial
    assert 0.2 < np.mean(g['contains']) < 0.8


def test_ellipsoid_known_answers():
    # tests/test_bounds.py:136-145: analytic volume of a sphere scaled by f.
    d = 10
    for f in (1.0, 1.1, np.pi / 2):
        ell = dict(c=np.full(d, 0.5), B=np.eye(d) * f, B_inv=np.eye(d) / f)
        assert np.isclose(orc.ell_log_v(ell), np.log(
            f**d * np.pi**(d / 2) / gamma(d / 2 + 1)))
    # tests/test_bounds.py:115-133, 148-156: samples are inside, radius < 1,
    # transform round-trips.
    rng = np.random.default_rng(0)
    ell = dict(c=np.full(d, 0.5), B=np.eye(d), B_inv=np.eye(d))
    pts = orc.ell_sample_from(ell, rng.normal(size=(100, d)),
                              rng.uniform(size=100))
    assert np.all(np.linalg.norm(pts - 0.5, axis=1) < 1 + 1e-9)
    assert np.all(orc.ell_contains(ell, pts))
    assert np.allclose(pts, orc.ell_transform(
        ell, orc.ell_transform(ell, pts), inverse=True))


def test_mixture_matches_reference(golden):
    g = golden('mixture_d6')
    spec = flat_to_spec(g)
    mix = spec['mixtures'][0]
    assert np.any(mix['dim_cube']) and not np.all(mix['dim_cube'])
    assert np.array_equal(orc.mix_contains(mix, g['points']), g['contains'])
    assert np.array_equal(orc.mix_transform(mix, g['points']), g['transform'])
    assert np.array_equal(
        orc.mix_sample_from(mix, g['cube_u'], g['z'], g['u']), g['sample'])
    assert orc.mix_log_v(mix) == g['log_v']


def test_union_matches_reference(golden):
    g = golden('union_d5')
    spec = flat_to_spec(g)
    assert len(spec['mixtures']) >= 3
    assert np.array_equal(orc.union_contains(spec, g['points']), g['contains'])
    assert np.array_equal(orc.union_count(spec, g['points']), g['count'])
    assert g['count'].max() >= 2          # overlaps are exercised
    # replay of one Union.sample iteration (bounds/union.py:305-323)
    raw = np.zeros_like(g['raw'])
    for k, mix in enumerate(spec['mixtures']):
        sel = g['k_assign'] == k
        de = int(np.sum(~mix['dim_cube']))
        nc = int(np.sum(mix['dim_cube']))
        raw[sel] = orc.mix_sample_from(
            mix, g['cube_u'][sel][:, :nc] if nc else None,
            g['z'][sel][:, :de], g['u'][sel])
    assert np.array_equal(raw, g['raw'])
    in_cube = orc.cube_contains(raw)
    assert np.array_equal(in_cube, g['in_cube'])
    shuffled = raw[in_cube][g['perm']]
    _, n_bound, accept = orc.union_accept(spec, shuffled, g['r'])
    assert np.array_equal(n_bound, g['n_bound'])
    assert np.array_equal(accept, g['accept'])
    assert np.array_equal(shuffled[accept], g['accepted'])
    assert int(g['n_reject']) == 1000 - int(np.sum(accept))
    assert orc.union_log_v(spec, int(g['n_sample']), int(g['n_reject'])) == \
        g['log_v']


@pytest.mark.parametrize('name', ['nautilus_d4', 'cfg2_bound_d30',
                                  'cfg5_bound_d100'])
def test_nautilus_bound_matches_reference(golden, name):
    g = golden(name)
    spec = flat_to_spec(g)
    nb = spec['neural'][0]
    pts = g['points']
    assert np.array_equal(orc.union_contains(spec, pts), g['union_contains'])
    assert np.array_equal(orc.ell_contains(nb['ell'], pts), g['ell_contains'])
    t = orc.ell_transform(nb['ell'], pts)
    pred = orc.emulator_predict(nb['emulator'], t)
    # sklearn's forward is x @ W + b through BLAS; the restatement issues the
    # same calls, so agreement is to the last bit on the same machine and to
    # rounding noise across BLAS builds.
    assert np.allclose(pred, g['predict'], rtol=0, atol=1e-13)
    assert np.array_equal(orc.neural_contains(nb, pts), g['neural_contains'])
    assert np.array_equal(orc.bound_contains(spec, pts), g['contains'])
    assert 0.02 < np.mean(g['contains']) < 0.98


def test_cfg2_volume_counters(golden):
    g = golden('cfg2_bound_d30')
    spec = flat_to_spec(g)
    log_v = orc.bound_log_v(spec, int(g['ref_u_n_sample']),
                            int(g['ref_u_n_reject']), int(g['ref_n_sample']),
                            int(g['ref_n_reject']))
    assert log_v == g['ref_log_v']


def test_shell_bookkeeping_matches_reference(golden):
    g = golden('shells_d2')
    n = int(g['n_shells'])
    assert n >= 3
    out = [orc.shell_info(g['log_l_{}'.format(i)], g['bound_log_v'][i],
                          g['shell_n_sample'][i]) for i in range(n)]
    shell_n = np.array([o[0] for o in out])
    shell_log_v = np.array([o[1] for o in out])
    shell_log_l = np.array([o[2] for o in out])
    shell_n_eff = np.array([o[3] for o in out])
    assert np.array_equal(shell_n, g['shell_n'])
    assert np.array_equal(shell_log_v, g['shell_log_v'])
    assert np.array_equal(shell_log_l, g['shell_log_l'])
    assert np.array_equal(shell_n_eff, g['shell_n_eff'])
    assert orc.log_z(shell_n, shell_log_l, shell_log_v) == g['log_z']
    assert orc.n_eff(shell_log_l, shell_log_v, shell_n_eff) == g['n_eff']
    assert orc.eta(shell_n, shell_log_l, shell_log_v, shell_n_eff) == g['eta']
    log_w = orc.posterior_log_w(
        shell_n, shell_log_v, [g['log_l_{}'.format(i)] for i in range(n)])
    assert np.array_equal(log_w, g['posterior_log_w'])
    # sampler.py:935-943 edge cases
    sn, lv, ll, ne = orc.shell_info(g['edge_all_minus_inf'], 0.0, 10)
    assert (sn, ll, ne) == (5, -np.inf, 5)
    sn, lv, ll, ne = orc.shell_info(np.zeros(0), 0.0, 10)
    assert sn == 0 and lv == -np.inf and np.isnan(ll) and ne == 0


def test_emulator_matches_reference(golden):
    g = golden('emulator_d5')
    emu = dict(mean=g['mean'], scale=g['scale'],
               coefs=[[g['W{}_{}'.format(n, i)] for i in range(4)]
                      for n in range(2)],
               intercepts=[[g['b{}_{}'.format(n, i)] for i in range(4)]
                           for n in range(2)])
    pred = orc.emulator_predict(emu, g['x'])
    assert np.allclose(pred, g['predict'], rtol=0, atol=1e-13)
    # tests/test_neural.py:15 quality bar of the reference
    assert np.sqrt(np.mean((g['y'] - pred)**2)) < 0.3 * np.std(g['y'])


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10 (kat_vectors):
    # ctr=0,key=0 ; ctr=ff..,key=ff.. ; ctr=pi digits, key=pi digits.
    out = philox.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(w) for w in out] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c,
                                     0x9b00dbd8]
    out = philox.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff,
                               0xffffffff, 0xffffffff)
    assert [int(w) for w in out] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6,
                                     0x6d5451fd]
    out = philox.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344,
                               0xa4093822, 0x299f31d0)
    assert [int(w) for w in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420,
                                     0x24126ea1]
    u = philox.u01_53(np.array([0xffffffff], np.uint32),
                      np.array([0xffffffff], np.uint32))
    assert u[0] < 1.0 and u[0] == 1 - 2.0**-53


def test_reference_arm_objects_reproduce_the_golden_bound(golden):
    """oracle/ref_arm.py rebuilds reference objects from the exported
    parameters (for the timed CPU arm): the rebuilt reference bound answers
    contains() exactly as the reference did when the fixture was made, and
    its own add_samples loop runs."""
    from oracle import ref_arm
    if not ref_arm.available():
        pytest.skip('oracle/_ref missing (run oracle/make_ref.sh)')
    from nautilus_b200._pack import flat_to_spec
    from nautilus_b200 import likelihoods
    g = golden('cfg2_bound_d30')
    spec = flat_to_spec(g)
    bound = ref_arm.reference_bound(spec, np.random.default_rng(0))
    assert np.array_equal(bound.contains(g['points']), g['contains'])
    like = likelihoods.Gaussian(30)
    n, dt, info = ref_arm.run_reference_cycles(
        spec, like, float(g['log_l_min']) + like.norm, 3000, seed=1,
        n_batch=100)
    assert n >= 3000 and info['shell_n'] >= 100 and info['n_like'] == \
        info['shell_n']
    assert np.isfinite(info['shell_log_l']) and info['shell_log_v'] < 0
