"""The host-buffer session (nb200_session_*): NumPy in, NumPy out, batches in
flight.  Results must be those of nb200_cycle + nb200_compact on the same
seeds, whatever the submit / wait interleaving, and agree with the oracle."""

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from nautilus_b200 import _lib, likelihoods, ops  # noqa: E402
from nautilus_b200._pack import flat_to_spec  # noqa: E402
from oracle import nautilus_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


def host(t):
    return t.cpu().numpy()


def shrink(spec, f):
    """A nested copy of a bound: every ellipsoid scaled by f about its
    centre (stands in for a later, smaller bound)."""
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


def _direct(spec, later, n, like, seed, offset, mode, log_l_min):
    stack = ops.DeviceStack([spec] + later)
    out = stack.cycle(0, n, later=(1, len(later)), seed=seed, offset=offset,
                      like_id=like.like_id,
                      like_params=like.device_params('cuda'),
                      log_l_min=log_l_min, mode=mode)
    sel = host(out['code']) == ops.CODE_IN_SHELL
    return (host(out['points'])[sel], host(out['log_l'])[sel],
            host(out['lse']), host(out['counters']))


@pytest.mark.parametrize('mode', [ops.MLP_F64, ops.MLP_TF32])
def test_session_matches_direct_cycle(golden, mode):
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    like = likelihoods.Gaussian(30)
    n = 1 << 14
    sess = ops.HostSession([spec], n_max=n, n_slots=2,
                           like_params_max=len(like.params()))
    # three batches, two in flight, waited out of submission order once
    offs = [0, n, 5 * n]
    sess.submit(0, 0, n, seed=3, offset=offs[0], like_id=like.like_id,
                like_params=like.params(), log_l_min=-30.0, mode=mode)
    sess.submit(1, 0, n, seed=3, offset=offs[1], like_id=like.like_id,
                like_params=like.params(), log_l_min=-30.0, mode=mode,
                upload_stack=True)
    got = [None, None, None]
    r = sess.wait(1)
    got[1] = {k: np.array(v) for k, v in r.items()}
    r = sess.wait(0)
    got[0] = {k: np.array(v) for k, v in r.items()}
    sess.submit(0, 0, n, seed=3, offset=offs[2], like_id=like.like_id,
                like_params=like.params(), log_l_min=-30.0, mode=mode)
    got[2] = {k: np.array(v) for k, v in sess.wait(0).items()}
    sess.close()
    for off, g in zip(offs, got):
        pts, ll, lse, cnt = _direct(spec, [], n, like, 3, off, mode, -30.0)
        assert np.array_equal(g['points'], pts)
        assert np.array_equal(g['log_l'], ll)
        assert np.array_equal(g['counters'], cnt)
        assert np.array_equal(g['lse'], lse)
        assert len(pts) == cnt[ops.CNT_IN_SHELL] > 0
    # and against the oracle: the likelihood of the returned rows, membership
    g = got[0]
    assert np.max(np.abs(g['log_l'] - like(g['points']))) < 1e-10
    if mode == ops.MLP_F64:
        assert np.all(orc.bound_contains(spec, g['points']))
        m, s1, s2 = orc.lse_triple(g['log_l'])
        assert g['lse'][0] == m and abs(g['lse'][1] / s1 - 1) < 1e-12


def test_session_staged_path_with_exclusion_and_new_stack(golden):
    spec = flat_to_spec(golden('nautilus_d4'))
    later = [shrink(spec, 0.9)]
    like = likelihoods.Gaussian(4, sigma=0.3)
    n = 20000
    sess = ops.HostSession([spec] + later, n_max=n, n_slots=1,
                           like_params_max=len(like.params()))
    sess.submit(0, 0, n, later=(1, 1), seed=9, like_id=like.like_id,
                like_params=like.params(), log_l_min=-1.0)
    g = sess.wait(0)
    pts, ll, lse, cnt = _direct(spec, later, n, like, 9, 0, ops.MLP_F64, -1.0)
    assert cnt[ops.CNT_EXCLUDED] > 0
    assert np.array_equal(g['points'], pts) and np.array_equal(g['log_l'], ll)
    assert np.array_equal(g['counters'], cnt)
    # a new (smaller-or-equal) stack replaces the old one: sampler.py:1023-1039
    sess.set_stack([later[0], spec])
    sess.submit(0, 0, n, seed=9, like_id=like.like_id,
                like_params=like.params(), log_l_min=-1.0)
    g2 = sess.wait(0)
    pts2, ll2, _, cnt2 = _direct(later[0], [], n, like, 9, 0, ops.MLP_F64,
                                 -1.0)
    assert np.array_equal(g2['points'], pts2)
    assert np.array_equal(g2['counters'], cnt2)
    # no likelihood: points only
    sess.submit(0, 0, 1000, seed=1)
    g3 = sess.wait(0)
    assert g3['log_l'] is None and len(g3['points']) == \
        g3['counters'][ops.CNT_IN_SHELL]
    sess.close()


def test_session_errors(golden):
    spec = flat_to_spec(golden('nautilus_d4'))
    like = likelihoods.Gaussian(4, sigma=0.3)
    sess = ops.HostSession([spec], n_max=4096, cap=8, n_slots=1,
                           like_params_max=len(like.params()))
    with pytest.raises(_lib.NautilusB200Error, match='n_max'):
        sess.submit(0, 0, 5000)
    with pytest.raises(_lib.NautilusB200Error, match='slot'):
        sess.submit(1, 0, 100)
    with pytest.raises(_lib.NautilusB200Error, match='nothing was submitted'):
        sess.wait(0)
    sess.submit(0, 0, 4096, seed=0)
    with pytest.raises(_lib.NautilusB200Error, match='un-waited'):
        sess.submit(0, 0, 4096, seed=1)
    with pytest.raises(_lib.NautilusB200Error, match='capacity'):
        sess.wait(0)          # more than cap=8 rows are in the shell
    # the slot is free again after the failed wait
    sess.submit(0, 0, 8, seed=0)
    assert len(sess.wait(0)['points']) <= 8
    with pytest.raises(_lib.NautilusB200Error):
        ops.HostSession([spec], n_max=0)
    sess.close()
    sess.close()              # idempotent


# --------------------------------------------------------------------------
# index mode: 16 bytes per in-shell proposal cross PCIe, rows are regenerated
# --------------------------------------------------------------------------

@pytest.mark.parametrize('mode,front', [(ops.MLP_F64, ''), (ops.MLP_TF32, ''),
                                        (ops.MLP_TF32, 'dfma'),
                                        (ops.MLP_F16, '')])
def test_index_session_and_materialize_bit_identical(golden, monkeypatch,
                                                     mode, front):
    """(index, log_l) of the index-mode session == the row-mode session's
    rows: log_l equal, materialised rows BIT-identical to the rows the cycle
    wrote, for each of the three proposal kernels (staged k_union_propose,
    DMMA front, DFMA front)."""
    if front:
        monkeypatch.setenv('NB200_FRONT', front)
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    like = likelihoods.Gaussian(30)
    n, seed, sid = 1 << 14, 11, 5
    kw = dict(seed=seed, stream_id=sid, like_id=like.like_id,
              like_params=like.params(), log_l_min=-30.0, mode=mode)
    rows = ops.HostSession([spec], n_max=n, n_slots=2,
                           like_params_max=len(like.params()))
    idx = ops.HostSession([spec], n_max=n, n_slots=2, returns='index',
                          like_params_max=len(like.params()))
    offs = [0, 3 * n, (1 << 40) + 7]       # indices beyond 32 bits too
    for slot, off in enumerate(offs):
        rows.submit(slot & 1, 0, n, offset=off, **kw)
        idx.submit(slot & 1, 0, n, offset=off, **kw)
        a = {k: np.array(v) for k, v in rows.wait(slot & 1).items()}
        b = {k: np.array(v) for k, v in idx.wait(slot & 1).items()}
        assert 'points' not in b and b['index'].dtype == np.uint64
        assert np.array_equal(a['log_l'], b['log_l'])
        assert np.array_equal(a['counters'], b['counters'])
        assert np.array_equal(a['lse'], b['lse'])
        assert len(b['index']) == a['counters'][ops.CNT_IN_SHELL] > 0
        assert np.all(np.diff(b['index'].astype(np.int64)) > 0)   # stable
        assert b['index'][0] >= off and b['index'][-1] < off + n
        got = idx.materialize(0, b['index'], seed=seed, stream_id=sid,
                              mode=mode)
        assert np.array_equal(got, a['points'])
        # any subset, any order
        pick = b['index'][::-7][:100]
        sub = idx.materialize(0, pick, seed=seed, stream_id=sid, mode=mode)
        assert np.array_equal(sub, a['points'][::-7][:100])
    # the first batch sized its copy from cap, the later ones from the
    # previous count: a batch that keeps far more than its predecessor still
    # comes back complete (top-up path)
    idx.submit(0, 0, 512, offset=0, **kw)
    few = idx.wait(0)
    assert len(few['index']) < 512
    idx.submit(0, 0, n, offset=0, **kw)
    many = {k: np.array(v) for k, v in idx.wait(0).items()}
    rows.submit(0, 0, n, offset=0, **kw)
    ref = rows.wait(0)
    assert np.array_equal(many['log_l'], ref['log_l'])
    assert len(many['index']) > len(few['index']) + len(few['index']) // 4 + \
        1024
    # the row-mode wait refuses a batch submitted in index mode
    idx.submit(0, 0, 100, **kw)
    with pytest.raises(_lib.NautilusB200Error, match='index mode'):
        _lib.check(_lib.lib().nb200_session_wait(idx._h, 0, None, None, None,
                                                 None, None))
    idx.wait(0)
    rows.close()
    idx.close()


def test_device_compact_index_and_materialize(golden):
    """Device-pointer forms: nb200_compact_index / nb200_materialize on a
    staged-path bound with cube dimensions and several ellipsoids."""
    spec = flat_to_spec(golden('nautilus_d4'))
    stack = ops.DeviceStack([spec])
    like = likelihoods.Gaussian(4, sigma=0.3)
    n, off = 30000, 123456789012
    out = stack.cycle(0, n, seed=4, offset=off, stream_id=2,
                      like_id=like.like_id,
                      like_params=like.device_params('cuda'))
    index, ll, n_out = stack.compact_index(out['log_l'], out['code'],
                                           offset=off)
    k = int(n_out.item())
    sel = host(out['code']) == ops.CODE_IN_SHELL
    assert k == sel.sum() > 0
    assert np.array_equal(host(index)[:k], off + np.flatnonzero(sel))
    assert np.array_equal(host(ll)[:k], host(out['log_l'])[sel])
    rows = stack.materialize(0, index[:k].contiguous(), seed=4, stream_id=2)
    assert np.array_equal(host(rows), host(out['points'])[sel])
    # rejected proposals can be regenerated too (rows are written for all)
    rej = torch.as_tensor(off + np.flatnonzero(~sel)[:50], device='cuda')
    assert np.array_equal(host(stack.materialize(0, rej, seed=4,
                                                 stream_id=2)),
                          host(out['points'])[~sel][:50])
    # empty
    e = stack.materialize(0, index[:0].contiguous(), seed=4, stream_id=2)
    assert e.shape == (0, 4)
    # the unit cube of shell 0
    cube = ops.DeviceStack([dict(kind='cube', n_dim=5)])
    o = cube.cycle(0, 1000, seed=1, offset=17)
    idx = torch.arange(17, 1017, device='cuda')
    assert np.array_equal(host(cube.materialize(0, idx, seed=1)),
                          host(o['points']))
