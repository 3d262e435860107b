"""Later-bound exclusion (HOT LOOP D, nautilus/sampler.py:796-801) as one
grouped pass: identical to the per-bound loop it replaces, consistent with the
oracle, launch count independent of the number of later bounds."""

import copy

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from nautilus_b200 import likelihoods, ops  # noqa: E402
from nautilus_b200._pack import flat_to_spec  # noqa: E402
from oracle import nautilus_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu

TOL_SCORE = 2e-3


def host(t):
    return t.cpu().numpy()


def shrunk(spec, f, shift=0.0, dthr=0.0):
    """A later, smaller bound: every ellipsoid scaled by f about its centre
    (moved by `shift`), the emulator threshold moved by `dthr`."""
    s = copy.deepcopy(spec)
    for m in s['mixtures']:
        if m['ell'] is not None:
            m['ell']['B'] = m['ell']['B'] * f
            m['ell']['B_inv'] = m['ell']['B_inv'] / f
            m['ell']['c'] = m['ell']['c'] + shift
    for nb in s['neural']:
        nb['ell']['B'] = nb['ell']['B'] * f
        nb['ell']['B_inv'] = nb['ell']['B_inv'] / f
        nb['ell']['c'] = nb['ell']['c'] + shift
        nb['score_predict_min'] = nb['score_predict_min'] + dthr
    return s


def later_bounds(spec, L):
    rng = np.random.default_rng(L)
    # (30 dimensions: a radius 3 % smaller is a volume 60 % smaller)
    return [shrunk(spec, 1.0 - 0.03 * (i + 1) / L,
                   shift=2e-4 * rng.normal(), dthr=0.02 * rng.normal())
            for i in range(L)]


def run(spec, later, n, like, seed, monkeypatch, which, chunk=None,
        mode=ops.MLP_TF32):
    if which == 'loop':
        monkeypatch.setenv('NB200_EXCLUDE', 'loop')
    else:
        monkeypatch.delenv('NB200_EXCLUDE', raising=False)
    if chunk:
        monkeypatch.setenv('NB200_EXCL_CHUNK', str(chunk))
    else:
        monkeypatch.delenv('NB200_EXCL_CHUNK', raising=False)
    stack = ops.DeviceStack([spec] + later)
    # warm-up call so that launch counts exclude one-time work
    stack.cycle(0, 256, later=(1, len(later)), seed=1, like_id=like.like_id,
                like_params=like.device_params('cuda'), mode=mode)
    torch.cuda.synchronize()
    l0 = ops.launch_count()
    out = stack.cycle(0, n, later=(1, len(later)), seed=seed, offset=11,
                      stream_id=3, like_id=like.like_id,
                      like_params=like.device_params('cuda'), log_l_min=-30.0,
                      mode=mode)
    torch.cuda.synchronize()
    launches = ops.launch_count() - l0
    return {k: host(v) for k, v in out.items()}, launches


@pytest.mark.parametrize('L', [1, 8, 47])
def test_grouped_exclusion_equals_loop_and_oracle(golden, monkeypatch, L):
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    like = likelihoods.Gaussian(30)
    later = later_bounds(spec, L)
    n = 1 << 15
    a, la = run(spec, later, n, like, 5, monkeypatch, 'grouped')
    b, lb = run(spec, later, n, like, 5, monkeypatch, 'loop')
    # bit-identical to the per-bound loop (same fp64 arithmetic, same tf32
    # emulator arithmetic row by row)
    assert np.array_equal(a['code'], b['code'])
    assert np.array_equal(a['points'], b['points'])
    assert np.array_equal(a['counters'], b['counters'])
    sel = a['code'] == ops.CODE_IN_SHELL
    assert np.array_equal(a['log_l'][sel], b['log_l'][sel])
    assert np.array_equal(a['lse'], b['lse'])
    assert a['counters'][ops.CNT_EXCLUDED] > 0 and sel.sum() > 0
    print('L = {}: grouped {} launches, loop {} launches; {} excluded, {} in '
          'shell'.format(L, la, lb, a['counters'][ops.CNT_EXCLUDED],
                         sel.sum()))
    assert la < 40
    # the oracle: dispositions differ only where some emulator score is
    # within the tf32 tolerance of its threshold
    _, r, _ = orc.replay_integer_stream(n, 11, 3, 5, spec)
    ref_code, _, ref_ll = orc.classify(spec, later, a['points'], r, like)
    bad = np.flatnonzero(ref_code != a['code'])
    assert len(bad) <= max(3, 2e-3 * n), len(bad)
    for i in bad:
        near = False
        for s in [spec] + later:
            nbs = s['neural'][0]
            _, score = orc.neural_contains(nbs, a['points'][i:i + 1],
                                           return_score=True)
            if np.isfinite(score[0]) and abs(
                    score[0] - (nbs['score_predict_min'] - 1e-9)) < TOL_SCORE:
                near = True
        assert near, (i, ref_code[i], a['code'][i])
    both = sel & (ref_code == ops.CODE_IN_SHELL)
    assert np.max(np.abs(a['log_l'][both] - ref_ll[both])) < 1e-10


def test_grouped_exclusion_launches_do_not_depend_on_L(golden, monkeypatch):
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    like = likelihoods.Gaussian(30)
    counts = {}
    for L in (2, 8, 47):
        _, counts[L] = run(spec, later_bounds(spec, L), 1 << 14, like, 2,
                           monkeypatch, 'grouped')
    assert counts[2] == counts[8] == counts[47], counts
    _, loop47 = run(spec, later_bounds(spec, 47), 1 << 14, like, 2,
                    monkeypatch, 'loop')
    assert loop47 > 4 * counts[47]


def test_grouped_exclusion_fp16_emulator(golden, monkeypatch):
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    like = likelihoods.Gaussian(30)
    later = later_bounds(spec, 8)
    a, la = run(spec, later, 1 << 15, like, 9, monkeypatch, 'grouped',
                mode=ops.MLP_F16)
    b, lb = run(spec, later, 1 << 15, like, 9, monkeypatch, 'loop',
                mode=ops.MLP_F16)
    assert np.array_equal(a['code'], b['code'])
    assert np.array_equal(a['counters'], b['counters'])
    assert a['counters'][ops.CNT_EXCLUDED] > 0 and la < lb
    c, _ = run(spec, later, 1 << 15, like, 9, monkeypatch, 'grouped')
    assert np.mean(a['code'] != c['code']) < 2e-3      # fp16 vs tf32


def test_grouped_exclusion_in_several_passes(golden, monkeypatch):
    # scratch for 1024 candidates per pass: ~5 passes over the candidates
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    like = likelihoods.Gaussian(30)
    later = later_bounds(spec, 8)
    a, _ = run(spec, later, 1 << 15, like, 7, monkeypatch, 'grouped',
               chunk=1024)
    b, _ = run(spec, later, 1 << 15, like, 7, monkeypatch, 'loop')
    assert np.array_equal(a['code'], b['code'])
    assert np.array_equal(a['counters'], b['counters'])


def test_grouped_exclusion_mixed_later_bounds(golden, monkeypatch):
    # later bounds of different make: several ellipsoids with cube
    # dimensions, a neural bound without emulator (ellipsoid only), and a
    # bound without neural bounds; staged front end (d = 4 fixture)
    spec = flat_to_spec(golden('nautilus_d4'))
    like = likelihoods.Gaussian(4, sigma=0.3)
    l1 = shrunk(spec, 0.9)
    l2 = shrunk(spec, 0.8, shift=0.01)
    l2['neural'][0]['emulator'] = None
    l2['neural'][0]['score_predict_min'] = 0
    l3 = shrunk(spec, 0.5)
    l3['neural'] = []
    second = copy.deepcopy(l1['mixtures'][0])
    second['ell']['c'] = second['ell']['c'] + 0.1
    l1['mixtures'].append(second)
    l1['log_v_all'] = np.repeat(l1['log_v_all'], 2)
    later = [l1, l2, l3]
    n = 20000
    a, _ = run(spec, later, n, like, 3, monkeypatch, 'grouped')
    b, _ = run(spec, later, n, like, 3, monkeypatch, 'loop')
    assert np.array_equal(a['code'], b['code'])
    assert np.array_equal(a['counters'], b['counters'])
    assert a['counters'][ops.CNT_EXCLUDED] > 0
    _, r, _ = orc.replay_integer_stream(n, 11, 3, 3, spec)
    ref_code, _, _ = orc.classify(spec, later, a['points'], r, like)
    assert np.mean(ref_code != a['code']) < 2e-3
