"""Parity of the CUDA path (through the C ABI) against the oracle.

Every test here needs a B200; they call ``libnautilus_b200.so`` via
``nautilus_b200.ops`` and compare with (a) the golden outputs of the real
reference, (b) the NumPy oracle on the same seeded inputs, (c) the
canonical-order C oracle bit-for-bit.  Bars: booleans / integers / indices
bit-exact; fp64 values bit-exact against the C oracle and within the stated
absolute tolerance of the NumPy reference path (summation order only).
"""

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from nautilus_b200 import likelihoods, ops  # noqa: E402
from nautilus_b200._pack import flat_to_spec, pack_stack  # noqa: E402
from oracle import c_oracle  # noqa: E402
from oracle import nautilus_oracle as orc  # noqa: E402
from oracle import philox  # noqa: E402

pytestmark = pytest.mark.gpu

# |numpy-einsum order - FMA-chain order| for O(1) values in <= 128 dims
TOL_F64 = 1e-13


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def host(t):
    return t.cpu().numpy()


def shrink(spec, f):
    """A nested copy of a bound: every ellipsoid scaled by f about its centre
    (stands in for a later, smaller bound in exclusion tests)."""
    import copy
    s = copy.deepcopy(spec)
    for m in s['mixtures']:
        if m['ell'] is not None:
            m['ell']['B'] = m['ell']['B'] * f
            m['ell']['B_inv'] = m['ell']['B_inv'] / f
    for nb in s['neural']:
        nb['ell']['B'] = nb['ell']['B'] * f
        nb['ell']['B_inv'] = nb['ell']['B_inv'] / f
    return s


# --------------------------------------------------------------------------
# a3-a6 Ellipsoid
# --------------------------------------------------------------------------

@pytest.mark.parametrize('d', [3, 10, 30])
def test_ellipsoid_golden(golden, d):
    g = golden('ellipsoid_d{}'.format(d))
    pts, c, B, Binv = dev(g['points']), dev(g['c']), dev(g['B']), dev(g['B_inv'])
    t = host(ops.ell_transform(pts, c, Binv))
    assert np.max(np.abs(t - g['transform'])) < TOL_F64
    assert np.array_equal(t, c_oracle.ell_transform(g['points'], g['c'],
                                                    g['B_inv']))
    inside, r2 = ops.ell_contains(pts, c, Binv, return_r2=True)
    assert np.array_equal(host(inside), g['contains'])         # bit-exact
    ref_in, ref_r2 = c_oracle.ell_contains(g['points'], g['c'], g['B_inv'])
    assert np.array_equal(host(r2), ref_r2)                    # bit-exact f64
    # safety margin of the boolean parity (SURVEY.md 7 "hard parts")
    assert np.min(np.abs(host(r2) - 1)) > 100 * TOL_F64
    back = host(ops.ell_transform(dev(g['transform']), c, B, inverse=True))
    assert np.max(np.abs(back - g['inverse'])) < TOL_F64
    assert np.array_equal(back, c_oracle.ell_transform(
        g['transform'], g['c'], g['B'], inverse=True))
    smp = host(ops.ell_sample_from(dev(g['z']), dev(g['u']), c, B))
    assert np.max(np.abs(smp - g['sample'])) < TOL_F64
    assert np.all(host(ops.ell_contains(dev(smp), c, Binv)))


@pytest.mark.parametrize('n', [0, 1, 127, 129])
def test_ellipsoid_ragged_sizes(n):
    d = 7
    rng = np.random.default_rng(n)
    B = np.tril(rng.normal(size=(d, d))) * 0.1 + np.eye(d)
    Binv = np.linalg.inv(B)
    c = rng.random(d)
    pts = rng.random((n, d)) * 2
    inside = host(ops.ell_contains(dev(pts).reshape(n, d), dev(c), dev(Binv)))
    assert inside.shape == (n,)
    if n:
        ref, _ = c_oracle.ell_contains(pts, c, Binv)
        assert np.array_equal(inside, ref)
        assert np.array_equal(inside, orc.ell_contains(
            dict(c=c, B=B, B_inv=Binv), pts))


def test_ellipsoid_max_dim():
    d, n = 128, 300
    rng = np.random.default_rng(1)
    B = np.tril(rng.normal(size=(d, d))) * 0.02 + np.eye(d) * 0.3
    Binv = np.linalg.inv(B)
    c = np.full(d, 0.5)
    ell = dict(c=c, B=B, B_inv=Binv)
    pts = orc.ell_sample_from(ell, rng.normal(size=(n, d)), rng.uniform(size=n))
    pts = np.vstack([pts, c + (pts - c) * 1.0005])
    inside, r2 = ops.ell_contains(dev(pts), dev(c), dev(Binv), return_r2=True)
    assert np.array_equal(host(inside), orc.ell_contains(ell, pts))
    assert np.array_equal(host(r2), c_oracle.ell_contains(pts, c, Binv)[1])
    with pytest.raises(RuntimeError):
        ops.ell_contains(dev(np.zeros((2, 129))), dev(np.zeros(129)),
                         dev(np.eye(129)))


# --------------------------------------------------------------------------
# a7-a9 mixture / union
# --------------------------------------------------------------------------

def test_mixture_golden(golden):
    g = golden('mixture_d6')
    spec = flat_to_spec(g)
    stack = ops.DeviceStack([spec])
    count, cont = stack.union_count(0, dev(g['points']))
    # Union.contains adds the unit-cube cut (union.py:287-288)
    assert np.array_equal(host(cont), g['contains'] &
                          orc.cube_contains(g['points']))
    assert np.array_equal(host(count), g['contains'].astype(np.int32))
    mix = spec['mixtures'][0]
    de, nc = int(np.sum(~mix['dim_cube'])), int(np.sum(mix['dim_cube']))
    n, d = g['sample'].shape
    z = np.zeros((n, d)); z[:, :de] = g['z']
    cu = np.zeros((n, d)); cu[:, :nc] = g['cube_u']
    pts, code, nb = stack.propose(0, n, test=dict(
        k=dev(np.zeros(n, np.int32)), z=dev(z), cube_u=dev(cu), u=dev(g['u']),
        r=dev(np.full(n, 0.5))))
    assert np.max(np.abs(host(pts) - g['sample'])) < TOL_F64
    # cube dims are copied, not computed: exact
    assert np.array_equal(host(pts)[:, mix['dim_cube']],
                          g['sample'][:, mix['dim_cube']])
    assert np.all(host(code) == ops.CODE_IN_SHELL)
    assert np.all(host(nb) == 1)


def test_union_golden(golden):
    g = golden('union_d5')
    spec = flat_to_spec(g)
    stack = ops.DeviceStack([spec])
    meta, data = pack_stack([spec])
    count, cont = stack.union_count(0, dev(g['points']))
    assert np.array_equal(host(count), g['count'])             # bit-exact
    assert np.array_equal(host(cont), g['contains'])
    c_count, c_cont = c_oracle.union_count(meta, data, 0, g['points'])
    assert np.array_equal(host(count), c_count)
    # masked evaluation leaves inactive points at 0
    mask = np.arange(len(g['points'])) % 3 == 0
    count_m, cont_m = stack.union_count(0, dev(g['points']), mask=dev(mask))
    assert np.array_equal(host(count_m), np.where(mask, g['count'], 0))

    # replay of the reference's Union.sample iteration in test mode
    n = len(g['raw'])
    in_cube = g['in_cube']
    inv = np.empty_like(g['perm']); inv[g['perm']] = np.arange(len(g['perm']))
    r_raw = np.full(n, 0.5)
    r_raw[np.flatnonzero(in_cube)] = g['r'][inv]
    pts, code, nb = stack.propose(0, n, test=dict(
        k=dev(g['k_assign'].astype(np.int32)), z=dev(g['z']),
        cube_u=dev(g['cube_u']), u=dev(g['u']), r=dev(r_raw)))
    pts, code, nb = host(pts), host(code), host(nb)
    assert np.max(np.abs(pts - g['raw'])) < TOL_F64
    assert np.array_equal(code != ops.CODE_CUBE_REJECT, in_cube)
    assert np.array_equal(nb[in_cube][g['perm']], g['n_bound'])  # bit-exact
    acc = (code == ops.CODE_IN_SHELL)[in_cube][g['perm']]
    assert np.array_equal(acc, g['accept'])
    assert int(np.sum(code != ops.CODE_IN_SHELL)) == int(g['n_reject'])


# --------------------------------------------------------------------------
# a11, a12, a16 NeuralBound / emulator / NautilusBound
# --------------------------------------------------------------------------

@pytest.mark.parametrize('name', ['nautilus_d4', 'cfg2_bound_d30',
                                  'cfg5_bound_d100'])
def test_nautilus_bound_golden(golden, name):
    g = golden(name)
    spec = flat_to_spec(g)
    stack = ops.DeviceStack([spec])
    meta, data = pack_stack([spec])
    pts = dev(g['points'])
    assert np.array_equal(host(stack.contains(0, pts, which=1)),
                          g['union_contains'])
    assert np.array_equal(host(stack.contains(0, pts, which=2)),
                          g['neural_contains'])
    assert np.array_equal(host(stack.contains(0, pts)), g['contains'])
    # emulator on the whitened coordinates
    in_ell, t_rows, score, ok = c_oracle.neural(meta, data, 0, 0, g['points'])
    assert np.array_equal(in_ell, g['ell_contains'])
    pred = host(stack.mlp_predict(0, 0, dev(t_rows)))
    assert np.max(np.abs(pred - g['predict'])) < 1e-12      # vs sklearn/BLAS
    assert np.array_equal(pred[in_ell], score[in_ell])      # vs C oracle
    # masked contains
    mask = np.arange(len(g['points'])) % 2 == 0
    assert np.array_equal(host(stack.contains(0, pts, mask=dev(mask))),
                          g['contains'] & mask)


def test_bound_without_emulator(golden):
    # n_networks=0: NeuralBound is its ellipsoid (bounds/neural.py:72-75)
    g = golden('nautilus_d4')
    spec = flat_to_spec(g)
    spec['neural'][0]['emulator'] = None
    spec['neural'][0]['score_predict_min'] = 0
    stack = ops.DeviceStack([spec])
    got = host(stack.contains(0, dev(g['points'])))
    assert np.array_equal(got, orc.bound_contains(spec, g['points']))
    assert np.array_equal(got, g['union_contains'] & g['ell_contains'])


def test_cube_record():
    stack = ops.DeviceStack([dict(kind='cube', n_dim=5)])
    rng = np.random.default_rng(0)
    pts = rng.random((1000, 5)) * 1.2 - 0.1
    pts[0] = 0.0; pts[1] = 1.0
    assert np.array_equal(host(stack.contains(0, dev(pts))),
                          orc.cube_contains(pts))
    p, code, _ = stack.propose(0, 4096, seed=3)
    p = host(p)
    assert np.all((p >= 0) & (p < 1)) and np.all(host(code) == 4)
    assert abs(p.mean() - 0.5) < 0.01
    # integer-exact replay of the uniforms
    idx = np.arange(4096, dtype=np.uint64)
    w = philox.philox_block(idx, 1, 0, 3)
    assert np.array_equal(p[:, 0], philox.u01_53(w[0], w[1]))
    assert np.array_equal(p[:, 1], philox.u01_53(w[2], w[3]))
    w = philox.philox_block(idx, 3, 0, 3)
    assert np.array_equal(p[:, 4], philox.u01_53(w[0], w[1]))


# --------------------------------------------------------------------------
# a19-a21 shell sums
# --------------------------------------------------------------------------

def test_stats_golden(golden):
    from scipy.special import logsumexp
    g = golden('shells_d2')
    for i in range(int(g['n_shells'])):
        ll = g['log_l_{}'.format(i)]
        lse, cnt = ops.stats(dev(ll), log_l_min=float(g['shell_log_l_min'][i]))
        m, s1, s2 = host(lse)[:3]
        cnt = host(cnt)
        assert cnt[ops.CNT_IN_SHELL] == len(ll) == cnt[ops.CNT_RAW]
        assert cnt[ops.CNT_UPDATE] == np.sum(ll >= g['shell_log_l_min'][i])
        shell_log_l = m + np.log(s1) - np.log(len(ll))
        n_eff = s1 * s1 / s2
        assert abs(shell_log_l - g['shell_log_l'][i]) < 1e-12
        assert abs(n_eff / g['shell_n_eff'][i] - 1) < 1e-12
        assert abs(m + np.log(s1) - logsumexp(ll)) < 1e-12


def test_stats_edge_cases():
    # all -inf, empty, single, huge dynamic range, mixed codes
    lse, cnt = ops.stats(dev(np.full(5, -np.inf)))
    assert host(lse)[0] == -np.inf and host(lse)[1] == 0
    assert host(cnt)[ops.CNT_IN_SHELL] == 5
    lse, cnt = ops.stats(dev(np.zeros(0)))
    assert host(cnt)[ops.CNT_RAW] == 0 and host(lse)[1] == 0
    lse, _ = ops.stats(dev(np.array([-3.5])))
    assert tuple(host(lse)[:3]) == (-3.5, 1.0, 1.0)
    rng = np.random.default_rng(0)
    ll = rng.normal(size=100003) * 300
    ll[::7] = -np.inf
    code = rng.integers(0, 5, size=len(ll)).astype(np.uint8)
    lse, cnt = ops.stats(dev(ll), code=dev(code), log_l_min=10.0)
    m, s1, s2 = orc.lse_triple(ll[code == 4])
    got = host(lse)
    assert got[0] == m
    assert abs(got[1] / s1 - 1) < 1e-12 and abs(got[2] / s2 - 1) < 1e-12
    cnt = host(cnt)
    assert cnt[ops.CNT_RAW] == len(ll)
    for c in range(4):
        assert cnt[1 + c] == np.sum(code == c)
    assert cnt[ops.CNT_IN_SHELL] == np.sum(code == 4)
    assert cnt[ops.CNT_UPDATE] == np.sum(ll[code == 4] >= 10.0)


def test_likelihoods_two_faces():
    rng = np.random.default_rng(0)
    for like in [likelihoods.Gaussian(30), likelihoods.Rosenbrock(50),
                 likelihoods.GaussianMixture(
                     0.5 + 0.25 * rng.choice([-1, 1], size=(4, 30)), 0.03),
                 likelihoods.EquicorrelatedGaussian(100)]:
        x = rng.random((1000, like.n_dim))
        got = host(ops.loglike(dev(x), like.like_id, like.device_params('cuda')))
        ref = like(x)
        assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref))) < 1e-12


# --------------------------------------------------------------------------
# a15, a17-a20 the cycle, production (Philox) mode
# --------------------------------------------------------------------------

def _check_cycle(spec, later_specs, n, like, seed, offset, stream_id,
                 log_l_min):
    stack = ops.DeviceStack([spec] + later_specs)
    out = stack.cycle(0, n, later=(1, len(later_specs)), seed=seed,
                      offset=offset, stream_id=stream_id,
                      like_id=like.like_id,
                      like_params=like.device_params('cuda'),
                      log_l_min=log_l_min)
    pts, code = host(out['points']), host(out['code'])
    log_l, lse, cnt = host(out['log_l']), host(out['lse']), host(out['counters'])
    # oracle decides every proposal's fate from the kernel's own raw draws
    k, r, u = orc.replay_integer_stream(n, offset, stream_id, seed, spec)
    ref_code, ref_nb, ref_ll = orc.classify(spec, later_specs, pts, r, like)
    assert np.array_equal(code, ref_code)                      # bit-exact
    sel = code == ops.CODE_IN_SHELL
    assert np.all(np.isnan(log_l[~sel]))
    assert np.max(np.abs(log_l[sel] - ref_ll[sel]) /
                  np.maximum(1, np.abs(ref_ll[sel]))) < 1e-12
    assert cnt[ops.CNT_RAW] == n
    for c in range(4):
        assert cnt[1 + c] == np.sum(ref_code == c)
    assert cnt[ops.CNT_IN_SHELL] == np.sum(sel)
    assert cnt[ops.CNT_UPDATE] == np.sum(log_l[sel] >= log_l_min)
    m, s1, s2 = orc.lse_triple(log_l[sel])
    assert lse[0] == m and abs(lse[1] / s1 - 1) < 1e-12
    assert abs(lse[2] / s2 - 1) < 1e-12
    # compaction is stable
    cp, cl, cn = stack.compact(out['points'], out['log_l'], out['code'])
    cn = int(cn.item())
    assert cn == np.sum(sel)
    assert np.array_equal(host(cp)[:cn], pts[sel])
    assert np.array_equal(host(cl)[:cn], log_l[sel])
    return pts, code, cnt


def test_cycle_cfg2(golden):
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    like = likelihoods.Gaussian(30)
    pts, code, cnt = _check_cycle(spec, [], 1 << 15, like, seed=7, offset=12345,
                                  stream_id=3, log_l_min=-20.0)
    # every proposal is a uniform draw from the ellipsoid: r^d ~ U(0,1)
    ell = spec['mixtures'][0]['ell']
    rd = np.sum(orc.ell_transform(ell, pts)**2, axis=1)**15
    from scipy.stats import kstest
    assert kstest(rd, 'uniform').pvalue > 1e-3
    t = orc.ell_transform(ell, pts)
    t /= np.linalg.norm(t, axis=1)[:, None]
    assert np.max(np.abs(t.mean(axis=0))) < 5 / np.sqrt(len(t) * 30)


def test_cycle_with_exclusion(golden):
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    later = [shrink(spec, 0.995), shrink(spec, 0.99)]
    like = likelihoods.Gaussian(30)
    pts, code, cnt = _check_cycle(spec, later, 1 << 14, like, seed=1, offset=0,
                                  stream_id=0, log_l_min=0.0)
    assert cnt[ops.CNT_EXCLUDED] > 0 and cnt[ops.CNT_IN_SHELL] > 0


def test_cycle_multi_ellipsoid(golden):
    g = golden('union_d5')
    spec = flat_to_spec(g)
    like = likelihoods.Gaussian(5, sigma=0.2)
    pts, code, cnt = _check_cycle(spec, [], 1 << 15, like, seed=11, offset=99,
                                  stream_id=1, log_l_min=-1.0)
    assert cnt[ops.CNT_OVERLAP_REJECT] > 0
    # volume from reject counts (union.py:342-343) against a plain Monte-Carlo
    # estimate of the same union volume from the oracle
    rng = np.random.default_rng(0)
    cube = rng.random((400000, 5))
    v_mc = np.mean(orc.union_contains(spec, cube))
    n_rej = cnt[ops.CNT_CUBE_REJECT] + cnt[ops.CNT_OVERLAP_REJECT]
    v_cnt = np.exp(orc.union_log_v(spec, int(cnt[ops.CNT_RAW]), int(n_rej)))
    assert abs(v_cnt / v_mc - 1) < 0.05


def test_cycle_shell0_cube():
    spec = dict(kind='cube', n_dim=4)
    like = likelihoods.Gaussian(4, sigma=0.2)
    _check_cycle(spec, [], 5000, like, seed=2, offset=0, stream_id=0,
                 log_l_min=-3.0)


def test_sharding_invariance(golden):
    # Philox is keyed by the global proposal index: splitting a batch over
    # launches / GPUs cannot change any proposal (cf. tests/test_bounds.py:
    # 412-441 of the reference, pool n_jobs in {1, 2}).
    spec = flat_to_spec(golden('union_d5'))
    stack = ops.DeviceStack([spec])
    n = 10000
    whole = stack.propose(0, n, seed=5, offset=1000, stream_id=2)
    a = stack.propose(0, 3333, seed=5, offset=1000, stream_id=2)
    b = stack.propose(0, n - 3333, seed=5, offset=1000 + 3333, stream_id=2)
    for w, x, y in zip(whole, a, b):
        assert torch.equal(w, torch.cat([x, y]))
    other = stack.propose(0, n, seed=6, offset=1000, stream_id=2)
    assert not torch.equal(whole[0], other[0])


def test_cycle_full_batch_properties(golden):
    # BASELINE config 2 size: 2^20 raw proposals through the whole cycle
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    stack = ops.DeviceStack([spec])
    like = likelihoods.Gaussian(30)
    n = 1 << 20
    out = stack.cycle(0, n, seed=0, like_id=like.like_id,
                      like_params=like.device_params('cuda'), log_l_min=-1e300)
    cnt = host(out['counters'])
    assert cnt[ops.CNT_RAW] == n and cnt[1:6].sum() == n
    cp, cl, cn = stack.compact(out['points'], out['log_l'], out['code'])
    cn = int(cn.item())
    assert cn == cnt[ops.CNT_IN_SHELL] == cnt[ops.CNT_UPDATE]
    # sample is a subset of contains (tests/test_bounds.py:330-349)
    assert bool(stack.contains(0, cp[:cn].contiguous()).all())
    # idempotence / determinism
    out2 = stack.cycle(0, n, seed=0, like_id=like.like_id,
                       like_params=like.device_params('cuda'),
                       log_l_min=-1e300)
    assert torch.equal(out['code'], out2['code'])
    assert torch.equal(out['lse'], out2['lse'])
    # LSE of the compacted log_l equals the fused one
    lse2, _ = ops.stats(cl[:cn].contiguous())
    assert abs(host(lse2)[1] / host(out['lse'])[1] - 1) < 1e-12


def test_cycle_host_entry(golden):
    import ctypes
    from nautilus_b200 import _lib
    spec = flat_to_spec(golden('nautilus_d4'))
    meta, data = pack_stack([spec])
    like = likelihoods.Gaussian(4, sigma=0.3)
    par = np.ascontiguousarray(like.params())
    n, cap, d = 20000, 20000, 4
    pts = np.empty((cap, d)); ll = np.empty(cap)
    n_out = ctypes.c_int64(); lse = np.empty(4); cnt = np.empty(8, np.int64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    _lib.check(_lib.lib().nb200_cycle_host(
        p(meta), len(meta), p(data), len(data), 0, 0, 0, n, 9, 0, 0,
        like.like_id, p(par), len(par), -1.0, ops.MLP_F64, cap, p(pts), p(ll),
        ctypes.byref(n_out), p(lse), p(cnt)))
    k = n_out.value
    assert k == cnt[ops.CNT_IN_SHELL] > 0
    assert np.all(orc.bound_contains(spec, pts[:k]))
    assert np.max(np.abs(ll[:k] - like(pts[:k]))) < 1e-12
    stack = ops.DeviceStack([spec])
    out = stack.cycle(0, n, seed=9, like_id=like.like_id,
                      like_params=like.device_params('cuda'), log_l_min=-1.0)
    assert np.array_equal(host(out['counters']), cnt)
    assert np.array_equal(host(out['points'])[host(out['code']) == 4], pts[:k])


def test_empty_and_degenerate_inputs(golden):
    # n = 0 everywhere, nothing accepted, one-point batches
    spec = flat_to_spec(golden('nautilus_d4'))
    stack = ops.DeviceStack([spec])
    empty = torch.empty((0, 4), dtype=torch.float64, device='cuda')
    assert stack.contains(0, empty).shape == (0,)
    assert stack.union_count(0, empty)[0].shape == (0,)
    p, code, nb = stack.propose(0, 0)
    assert p.shape == (0, 4) and code.numel() == 0
    like = likelihoods.Gaussian(4, sigma=0.3)
    out = stack.cycle(0, 0, like_id=like.like_id,
                      like_params=like.device_params('cuda'))
    assert host(out['counters'])[ops.CNT_RAW] == 0
    assert host(out['lse'])[0] == -np.inf
    # a batch in which nothing survives: every point far outside
    far = dev(np.full((100, 4), 5.0))
    assert not bool(stack.contains(0, far).any())
    code0 = torch.zeros(100, dtype=torch.uint8, device='cuda')
    cp, cl, cn = stack.compact(far, torch.zeros(100, dtype=torch.float64,
                                                device='cuda'), code0)
    assert int(cn.item()) == 0
    # wrong dimensionality / dtype are refused at the Python boundary
    with pytest.raises(ValueError):
        stack.contains(0, dev(np.zeros((3, 5))))
    with pytest.raises(ValueError):
        stack.contains(0, dev(np.zeros((3, 4), dtype=np.float32)))
    # out-of-range bound index is refused by the library
    with pytest.raises(RuntimeError):
        stack.contains(3, dev(np.zeros((3, 4))))


def test_torch_ops_match_the_ctypes_binding(golden):
    """torch.ops.nautilus_b200.* (stable-ABI registration) return exactly what
    the ctypes binding of the same C entry points returns."""
    from nautilus_b200 import _lib, likelihoods
    tops = _lib.torch_ops()
    g = golden('ellipsoid_d10')
    pts = dev(g['points'])
    got = tops.ell_contains(pts, dev(g['c']), dev(g['B_inv']))
    assert got.dtype == torch.uint8
    assert np.array_equal(host(got).astype(bool), g['contains'])
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    stack = ops.DeviceStack([spec])
    like = likelihoods.Gaussian(30)
    par = like.device_params('cuda')
    meta_h = torch.from_numpy(stack.meta_h)
    for mode in (ops.MLP_F64, ops.MLP_F16):
        ref = stack.cycle(0, 5000, seed=7, offset=3, stream_id=1,
                          like_id=like.like_id, like_params=par,
                          log_l_min=-20.0, mode=mode)
        p, ll, code, lse, cnt = tops.shell_cycle(
            meta_h, stack.meta_d, stack.data_d, 0, 0, 0, 5000, 7, 3, 1,
            like.like_id, par, -20.0, mode)
        assert torch.equal(p, ref['points']) and torch.equal(code, ref['code'])
        assert torch.equal(cnt, ref['counters']) and torch.equal(lse,
                                                                 ref['lse'])
        sel = code == ops.CODE_IN_SHELL
        assert torch.equal(ll[sel], ref['log_l'][sel])
        inside = tops.bound_contains(meta_h, stack.meta_d, stack.data_d, 0, 0,
                                     p, mode)
        assert torch.equal(inside.bool(), stack.contains(0, p, mode=mode))
        lse2, cnt2 = tops.shell_stats(ll[sel].contiguous(), -20.0)
        # (another reduction shape: same maximum, sums to rounding)
        assert float(lse2[0]) == float(lse[0])
        assert torch.allclose(lse2[1:3], lse[1:3], rtol=1e-12, atol=0)
        assert int(cnt2[ops.CNT_UPDATE]) == int(cnt[ops.CNT_UPDATE])
    with pytest.raises(RuntimeError, match='bound index'):
        tops.shell_cycle(meta_h, stack.meta_d, stack.data_d, 5, 0, 0, 10, 0,
                         0, 0, -1, par, 0.0, 0)


def test_device_radix_select_kth_largest():
    """nb200_select_kth_largest (the live-set threshold, sampler.py:1007-1009)
    against numpy's sort: every bit of the threshold, the count above it;
    ties, negative values, -inf, a mask, k = 1 and k = n."""
    rng = np.random.default_rng(0)
    cases = [rng.normal(size=100003) * 50,
             -np.abs(rng.normal(size=5000)) * 1e-300,
             np.round(rng.normal(size=20000), 1),          # many ties
             np.concatenate([np.full(100, -np.inf), rng.normal(size=900)]),
             np.array([3.0]), np.zeros(1000)]
    for v in cases:
        t = dev(v)
        srt = np.sort(v)[::-1]
        for k in sorted({1, len(v), max(1, len(v) // 3), min(len(v), 2000)}):
            thr, greater = ops.kth_largest(t, k)
            assert host(thr)[0] == srt[k - 1]
            assert int(host(greater)[0]) == np.sum(v > srt[k - 1])
            idx = host(ops.top_k(t, k))
            assert len(idx) == min(k, len(v)) and len(set(idx)) == len(idx)
            assert np.array_equal(np.sort(v[idx])[::-1], srt[:k])
    v = cases[0]
    mask = rng.random(len(v)) < 0.3
    thr, greater = ops.kth_largest(dev(v), 500, mask=dev(mask))
    assert host(thr)[0] == np.sort(v[mask])[::-1][499]
    with pytest.raises(_lib_error()):
        ops.kth_largest(dev(v), len(v) + 1)


def _lib_error():
    from nautilus_b200._lib import NautilusB200Error
    return NautilusB200Error
