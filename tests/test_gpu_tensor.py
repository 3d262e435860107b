"""tcgen05 emulator path (NB200_MLP_TF32) against the fp64 parity path.

tf32 x tf32 -> fp32 arithmetic cannot be bit-exact against scikit-learn's
fp64 forward pass; the bar is stated here: scores within TOL_SCORE = 2e-3
absolute of the fp64 kernel (measured 1.1e-3), NeuralBound membership flips
only for points whose fp64 score lies within that band of the threshold, and
fewer than TOL_FLIPS = 1e-3 of the proposals flip (measured 2.5e-4).
"""

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from nautilus_b200 import ops  # noqa: E402
from nautilus_b200._pack import flat_to_spec, pack_stack  # noqa: E402
from oracle import c_oracle  # noqa: E402

pytestmark = pytest.mark.gpu

TOL_SCORE = 2e-3
TOL_FLIPS = 1e-3


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


TC_MODES = [ops.MLP_TF32, ops.MLP_F16]
MODE_NAME = {ops.MLP_TF32: 'tf32', ops.MLP_F16: 'f16'}


@pytest.mark.parametrize('mode', TC_MODES)
@pytest.mark.parametrize('name', ['cfg2_bound_d30', 'nautilus_d4'])
def test_tf32_scores_close_to_fp64(golden, name, mode):
    g = golden(name)
    spec = flat_to_spec(g)
    stack = ops.DeviceStack([spec])
    meta, data = pack_stack([spec])
    in_ell, t_rows, score, ok = c_oracle.neural(meta, data, 0, 0, g['points'])
    t = dev(t_rows)
    p64 = stack.mlp_predict(0, 0, t, mode=ops.MLP_F64).cpu().numpy()
    p32 = stack.mlp_predict(0, 0, t, mode=mode).cpu().numpy()
    err = np.abs(p32 - p64)[in_ell]
    print('{} emulator: max |d score| = {:.2e}, mean = {:.2e}'.format(
        MODE_NAME[mode], err.max(), err.mean()))
    assert err.max() < TOL_SCORE
    assert np.max(np.abs(p32[in_ell] - g['predict'][in_ell])) < TOL_SCORE


@pytest.mark.parametrize('mode', TC_MODES)
def test_tf32_membership_flips_only_at_threshold(golden, mode):
    g = golden('cfg2_bound_d30')
    spec = flat_to_spec(g)
    stack = ops.DeviceStack([spec])
    rng = np.random.default_rng(0)
    n = 200000
    pts, code, _ = stack.propose(0, n, seed=4)
    c64 = stack.contains(0, pts, which=2, mode=ops.MLP_F64).cpu().numpy()
    c32 = stack.contains(0, pts, which=2, mode=mode).cpu().numpy()
    flips = c64 != c32
    rate = flips.mean()
    print('{} membership flip rate = {:.3e} ({} of {})'.format(
        MODE_NAME[mode], rate, flips.sum(), n))
    assert rate < TOL_FLIPS
    # every flip sits within the score tolerance of the threshold
    meta, data = pack_stack([spec])
    idx = np.flatnonzero(flips)
    if len(idx):
        _, _, score, _ = c_oracle.neural(meta, data, 0, 0,
                                         pts.cpu().numpy()[idx])
        thr = spec['neural'][0]['score_predict_min'] - 1e-9
        assert np.max(np.abs(score - thr)) < TOL_SCORE
    del rng


def test_tf32_ragged_and_masked(golden):
    g = golden('cfg2_bound_d30')
    spec = flat_to_spec(g)
    stack = ops.DeviceStack([spec])
    pts, _, _ = stack.propose(0, 1000, seed=1)
    ref = stack.contains(0, pts, which=2, mode=ops.MLP_TF32).cpu().numpy()
    for n in (1, 127, 129, 257, 999):
        got = stack.contains(0, pts[:n].contiguous(), which=2,
                             mode=ops.MLP_TF32).cpu().numpy()
        assert np.array_equal(got, ref[:n])
    mask = torch.arange(1000, device='cuda') % 3 == 0
    got = stack.contains(0, pts, which=2, mask=mask,
                         mode=ops.MLP_TF32).cpu().numpy()
    assert np.array_equal(got, ref & mask.cpu().numpy())


def test_tf32_cycle_self_consistent(golden):
    # the same kernel filters proposals and answers contains(): every point
    # the cycle keeps is inside the bound under the same arithmetic
    from nautilus_b200 import likelihoods
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    stack = ops.DeviceStack([spec])
    like = likelihoods.Gaussian(30)
    out = stack.cycle(0, 1 << 17, seed=3, like_id=like.like_id,
                      like_params=like.device_params('cuda'),
                      mode=ops.MLP_TF32)
    cp, cl, cn = stack.compact(out['points'], out['log_l'], out['code'])
    cn = int(cn.item())
    assert cn > 0
    assert bool(stack.contains(0, cp[:cn].contiguous(),
                               mode=ops.MLP_TF32).all())
    out64 = stack.cycle(0, 1 << 17, seed=3, like_id=like.like_id,
                        like_params=like.device_params('cuda'),
                        mode=ops.MLP_F64)
    diff = (out['code'] != out64['code']).float().mean().item()
    print('cycle disposition mismatch tf32 vs f64: {:.3e}'.format(diff))
    assert diff < TOL_FLIPS


def _fused_vs_staged(spec, n, like, seed, flip_tol=TOL_FLIPS,
                     mode=ops.MLP_TF32):
    """The fused tf32 cycle (k_front -> k_mlp_tf32 with likelihood and shell
    sums in its tail) against the same decisions rebuilt from staged ops."""
    from oracle import nautilus_oracle as orc
    stack = ops.DeviceStack([spec])
    out = stack.cycle(0, n, seed=seed, offset=77, stream_id=5,
                      like_id=like.like_id,
                      like_params=like.device_params('cuda'), log_l_min=-5.0,
                      mode=mode)
    pts = out['points']
    code = out['code'].cpu().numpy()
    log_l = out['log_l'].cpu().numpy()
    # union stage decided in fp64: bit-exact against the oracle
    _, r, _ = orc.replay_integer_stream(n, 77, 5, seed, spec)
    p = pts.cpu().numpy()
    in_cube = orc.cube_contains(p)
    nb = np.zeros(n, dtype=int)
    nb[in_cube] = orc.union_count(spec, p[in_cube])
    with np.errstate(divide='ignore'):
        acc = in_cube & (r > 1 - 1.0 / np.maximum(nb, 0))
    assert np.array_equal(code == 0, ~in_cube)
    assert np.array_equal(code == 1, in_cube & ~acc)
    # neural stage: same arithmetic as contains(which=2, tf32)
    nn = stack.contains(0, pts, which=2, mode=mode).cpu().numpy()
    assert np.array_equal(code == 4, acc & nn)
    assert np.array_equal(code == 2, acc & ~nn)
    # ... and against the fp64 ORACLE (not the same kernel): the dispositions
    # of orc.classify differ only where the oracle's own fp64 score lies
    # within TOL_SCORE of the threshold
    ref_code, _, _ = orc.classify(spec, [], p, r, None)
    bad = np.flatnonzero(ref_code != code)
    assert len(bad) <= max(2, flip_tol * n), (len(bad), n)
    if len(bad):
        assert set(ref_code[bad]) | set(code[bad]) <= {2, 4}
        nbs = spec['neural'][0]
        _, score = orc.neural_contains(nbs, p[bad], return_score=True)
        thr = nbs['score_predict_min'] - 1e-9
        assert np.nanmax(np.abs(score - thr)) < TOL_SCORE
    sel = code == 4
    ref_ll = like(p[sel])
    assert np.all(np.isnan(log_l[~sel]))
    if sel.any():
        assert np.max(np.abs(log_l[sel] - ref_ll) /
                      np.maximum(1, np.abs(ref_ll))) < 1e-12
    cnt = out['counters'].cpu().numpy()
    lse = out['lse'].cpu().numpy()
    assert cnt[ops.CNT_RAW] == n
    for c in range(4):
        assert cnt[1 + c] == np.sum(code == c)
    assert cnt[ops.CNT_IN_SHELL] == np.sum(sel)
    assert cnt[ops.CNT_UPDATE] == np.sum(log_l[sel] >= -5.0)
    m, s1, s2 = orc.lse_triple(log_l[sel])
    assert lse[0] == m
    if sel.any():
        assert abs(lse[1] / s1 - 1) < 1e-12 and abs(lse[2] / s2 - 1) < 1e-12
    return cnt


@pytest.mark.parametrize('mode', TC_MODES)
def test_fused_cycle_single_ellipsoid(golden, mode):
    from nautilus_b200 import likelihoods
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    for n in (1, 200, 1 << 15):
        _fused_vs_staged(spec, n, likelihoods.Gaussian(30), seed=n, mode=mode)


def test_fused_cycle_overlapping_ellipsoids(golden):
    # K = 2 overlapping ellipsoids, the neural bound shares mixture 0's
    import copy
    from nautilus_b200 import likelihoods
    spec = flat_to_spec(golden('nautilus_d4'))
    second = copy.deepcopy(spec['mixtures'][0])
    second['ell']['c'] = second['ell']['c'] + 0.15
    spec['mixtures'].append(second)
    spec['log_v_all'] = np.repeat(spec['log_v_all'], 2)
    cnt = _fused_vs_staged(spec, 1 << 15, likelihoods.Gaussian(4, sigma=0.3),
                           seed=3)
    assert cnt[ops.CNT_OVERLAP_REJECT] > 0 and cnt[ops.CNT_CUBE_REJECT] > 0
    # the DFMA front kernel writing fp16 emulator rows
    _fused_vs_staged(spec, 1 << 14, likelihoods.Gaussian(4, sigma=0.3),
                     seed=5, mode=ops.MLP_F16)
    # and with the neural ellipsoid different from every mixture
    spec['neural'][0]['ell']['c'] = spec['neural'][0]['ell']['c'] + 1e-3
    _fused_vs_staged(spec, 1 << 14, likelihoods.Gaussian(4, sigma=0.3), seed=4)


def test_fused_cycle_single_ellipsoid_distinct_neural_ellipsoid(golden):
    # K = 1 with the neural bound's ellipsoid different from the mixture's:
    # the DMMA front kernel runs two whitenings
    from nautilus_b200 import likelihoods
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    spec['neural'][0]['ell']['c'] = spec['neural'][0]['ell']['c'] + 1e-3
    _fused_vs_staged(spec, 1 << 14, likelihoods.Gaussian(30), seed=9)
    # no unit-cube cut
    spec['unit'] = False
    cnt = _fused_vs_staged(spec, 5000, likelihoods.Gaussian(30), seed=10)
    assert cnt[ops.CNT_CUBE_REJECT] == 0


def test_front_kernels_agree(golden, monkeypatch):
    # the DMMA front kernel (one ellipsoid) against the DFMA front kernel on
    # the same Philox streams: same proposals up to fp64 rounding (the DMMA
    # sums in another order and applies the radial factor after the product),
    # same dispositions
    from nautilus_b200 import likelihoods
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    like = likelihoods.Gaussian(30)
    n = 1 << 15
    outs = []
    for which in ('mma', 'dfma'):
        monkeypatch.setenv('NB200_FRONT', which)
        stack = ops.DeviceStack([spec])
        out = stack.cycle(0, n, seed=21, offset=5, stream_id=2,
                          like_id=like.like_id,
                          like_params=like.device_params('cuda'),
                          log_l_min=0.0, mode=ops.MLP_TF32)
        outs.append({k: v.cpu().numpy() for k, v in out.items()})
    a, b = outs
    assert np.max(np.abs(a['points'] - b['points'])) < 1e-13
    assert np.mean(a['code'] != b['code']) < 1e-4
    same = (a['code'] == 4) & (b['code'] == 4)
    assert np.max(np.abs(a['log_l'][same] - b['log_l'][same])) < 1e-9
    assert np.abs(a['counters'] - b['counters']).max() <= 3


@pytest.mark.parametrize('d', [3, 8, 17, 33, 40, 50, 64, 72, 81, 90, 100, 111,
                               120, 128])
def test_dmma_front_kernel_all_row_widths(d, monkeypatch):
    # every instantiation of k_front_mma<D8> (D8 = 8 ... 128; above 64 with
    # triangular-packed factors and fewer warps per CTA): a hand-built
    # one-ellipsoid bound with a small emulator; decisions against the oracle
    # on the kernel's own proposals, and agreement with the DFMA kernel
    from nautilus_b200 import bounds, likelihoods
    from nautilus_b200.neural import NeuralNetworkEmulator
    rng = np.random.default_rng(d)
    z = rng.normal(size=(600, d))
    z *= (rng.random((600, 1))**(1.0 / d) /
          np.linalg.norm(z, axis=1)[:, None])
    mix = np.eye(d) * 0.3 + 0.03 * rng.normal(size=(d, d))
    live = 0.5 + z @ mix.T
    ell = bounds.Ellipsoid.compute(live, enlarge_per_dim=1.05,
                                   rng=np.random.default_rng(0))
    whitened = ell.transform(live)
    score = 1.0 - np.linalg.norm(whitened, axis=1)
    emu = NeuralNetworkEmulator.train(
        whitened, score, n_networks=2,
        neural_network_kwargs=dict(hidden_layer_sizes=(16, 8), max_iter=15),
        seed=1)
    spec = dict(kind='nautilus', n_dim=d, unit=True,
                log_v_all=np.array([ell.log_v]),
                mixtures=[dict(dim_cube=np.zeros(d, dtype=bool),
                               ell=ell.ell_spec())],
                neural=[dict(ell=ell.ell_spec(), emulator=emu.emu_spec(),
                             score_predict_min=float(np.median(
                                 emu.predict(whitened))))])
    like = likelihoods.Gaussian(d, sigma=0.2)
    # (a 15-epoch toy network thresholded at its median score: many points
    # sit at the threshold, hence the wider flip allowance)
    cnt = _fused_vs_staged(spec, 6000, like, seed=d, flip_tol=5e-3)
    _fused_vs_staged(spec, 3000, like, seed=d + 1, flip_tol=5e-3,
                     mode=ops.MLP_F16)
    assert cnt[ops.CNT_IN_SHELL] > 0 and cnt[ops.CNT_NN_REJECT] > 0
    outs = []
    for which in ('mma', 'dfma'):
        monkeypatch.setenv('NB200_FRONT', which)
        stack = ops.DeviceStack([spec])
        out = stack.cycle(0, 4096, seed=3, like_id=like.like_id,
                          like_params=like.device_params('cuda'),
                          log_l_min=-1.0, mode=ops.MLP_TF32)
        outs.append({k: v.cpu().numpy() for k, v in out.items()})
    a, b = outs
    assert np.max(np.abs(a['points'] - b['points'])) < 1e-12
    assert np.mean(a['code'] != b['code']) < 2e-3


def test_fused_cycle_config5_bound(golden):
    """BASELINE config 5 (100-D, the bound the reference built): the wide DMMA
    front end + the fp16 emulator in its one-tile-group form + fused shell
    sums, against the oracle and the staged kernels."""
    from nautilus_b200 import likelihoods
    spec = flat_to_spec(golden('cfg5_bound_d100'))
    like = likelihoods.EquicorrelatedGaussian(100)
    stack = ops.DeviceStack([spec])
    l0 = ops.launch_count()
    stack.cycle(0, 4096, seed=1, like_id=like.like_id,
                like_params=like.device_params('cuda'), mode=ops.MLP_F16)
    assert ops.launch_count() - l0 == 3          # front, emulator, final sums
    cnt = _fused_vs_staged(spec, 1 << 15, like, seed=2, mode=ops.MLP_F16)
    assert cnt[ops.CNT_NN_REJECT] > 0
    _fused_vs_staged(spec, 3000, like, seed=3, mode=ops.MLP_TF32)


@pytest.mark.parametrize('name,n', [('cfg2_bound_d30', 1 << 17),
                                    ('nautilus_d4', 1 << 15),
                                    ('cfg5_bound_d100', 1 << 14)])
def test_front_whitening_shortcut_decides_like_the_exact_whitening(
        golden, monkeypatch, name, n):
    """k_front_mma<.., FAST>: when the neural bound's ellipsoid is the
    mixture's, the whitened point of x = c + s B z is taken as s z and the
    second DMMA pass is skipped unless r^2 is within a guard band of 1.
    NB200_FRONT_TAU overrides the band: 2 sends EVERY tile through the exact
    whitening (the kernel of the previous round), 1e-3 about half of them.
    Dispositions, likelihoods, sums and counters must not depend on it."""
    from nautilus_b200 import likelihoods
    spec = flat_to_spec(golden(name))
    d = int(spec['n_dim'])
    like = likelihoods.Gaussian(d, sigma=0.3)
    stack = ops.DeviceStack([spec])
    for mode in (ops.MLP_F16, ops.MLP_TF32):
        runs = []
        for tau in ('2', '1e-3', None):
            if tau is None:
                monkeypatch.delenv('NB200_FRONT_TAU', raising=False)
            else:
                monkeypatch.setenv('NB200_FRONT_TAU', tau)
            out = stack.cycle(0, n, seed=11, offset=123, stream_id=2,
                              like_id=like.like_id,
                              like_params=like.device_params('cuda'),
                              log_l_min=-5.0, mode=mode)
            runs.append({k: out[k].clone() for k in
                         ('points', 'code', 'log_l', 'counters', 'lse')})
        exact = runs[0]
        assert int((exact['code'] == ops.CODE_IN_SHELL).sum()) > 0
        for other in runs[1:]:
            assert torch.equal(exact['points'], other['points'])
            assert torch.equal(exact['code'], other['code'])
            assert torch.equal(exact['counters'], other['counters'])
            assert torch.equal(exact['log_l'].nan_to_num(nan=-7e77),
                               other['log_l'].nan_to_num(nan=-7e77))
            assert torch.allclose(exact['lse'], other['lse'], rtol=1e-13,
                                  atol=0)


def _synthetic_spec(d, rng, hidden=(24, 12)):
    """One ellipsoid shared by the mixture and the neural bound, two small
    random networks."""
    m = rng.normal(size=(d, d))
    cov = 0.004 * (m @ m.T / d + np.eye(d))
    B = np.linalg.cholesky(cov)
    ell = dict(c=0.5 + 0.01 * rng.normal(size=d), B=B,
               B_inv=np.tril(np.linalg.solve(B, np.eye(d))))
    sizes = (d, ) + hidden + (1, )
    coefs = [[rng.normal(size=(a, b)) / np.sqrt(a)
              for a, b in zip(sizes[:-1], sizes[1:])] for _ in range(2)]
    intercepts = [[0.1 * rng.normal(size=b) for b in sizes[1:]]
                  for _ in range(2)]
    emu = dict(mean=0.05 * rng.normal(size=d),
               scale=0.5 + rng.uniform(size=d), coefs=coefs,
               intercepts=intercepts)
    return dict(kind='nautilus', n_dim=d, unit=True, log_v_all=np.zeros(1),
                mixtures=[dict(dim_cube=np.zeros(d, bool), ell=ell)],
                neural=[dict(ell={k: v.copy() for k, v in ell.items()},
                             emulator=emu, score_predict_min=0.0)])


@pytest.mark.parametrize('d', [8, 16, 32, 47, 64])
def test_front_whitening_shortcut_row_widths(monkeypatch, d):
    """Row widths at and around multiples of 8 / 16: the bias column and the
    zero padding of the emulator's input row lie behind the padded row of the
    proposal (d = 16, 32, 64) or inside it; the shortcut must hand the
    emulator what the exact whitening hands it."""
    from nautilus_b200 import likelihoods
    spec = _synthetic_spec(d, np.random.default_rng(d))
    meta, _ = pack_stack([spec])
    rec = meta[meta[1]:]
    assert rec[10] == 1 and 0 <= rec[11] < 30
    like = likelihoods.Gaussian(d, sigma=0.2)
    stack = ops.DeviceStack([spec])
    n = 1 << 14
    for mode in (ops.MLP_F16, ops.MLP_TF32):
        runs = []
        for tau in ('2', None):
            if tau is None:
                monkeypatch.delenv('NB200_FRONT_TAU', raising=False)
            else:
                monkeypatch.setenv('NB200_FRONT_TAU', tau)
            out = stack.cycle(0, n, seed=d, offset=5, stream_id=1,
                              like_id=like.like_id,
                              like_params=like.device_params('cuda'),
                              log_l_min=-5.0, mode=mode)
            runs.append({k: out[k].clone() for k in
                         ('points', 'code', 'log_l', 'counters')})
        exact, fast = runs
        frac = float((exact['code'] == ops.CODE_IN_SHELL).float().mean())
        assert 1e-3 < frac < 0.999, frac
        assert torch.equal(exact['points'], fast['points'])
        assert torch.equal(exact['code'], fast['code'])
        assert torch.equal(exact['counters'], fast['counters'])
        assert torch.equal(exact['log_l'].nan_to_num(nan=-7e77),
                           fast['log_l'].nan_to_num(nan=-7e77))
        # the fused decisions are those of contains() on the finished rows
        monkeypatch.delenv('NB200_FRONT_TAU', raising=False)
        inside = stack.contains(0, fast['points'], mode=mode)
        assert torch.equal(inside, fast['code'] == ops.CODE_IN_SHELL)
