"""The BASELINE.json configurations beyond config 2, at sizes that run in
seconds: multi-modal targets (multi-ellipsoid union, several neural bounds),
a wider input that needs the single-tile-group mode of the tensor-core kernel,
and a 100-D correlated Gaussian on a fixed bound (fp64 emulator path)."""

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from nautilus_b200 import Sampler, bounds, likelihoods, ops  # noqa: E402
from nautilus_b200._pack import pack_stack  # noqa: E402
from oracle import nautilus_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


def _ball_points(rng, n, d, radius, centre=0.5):
    z = rng.normal(size=(n, d))
    z /= np.linalg.norm(z, axis=1)[:, None]
    return centre + radius * z * rng.random((n, 1))**(1.0 / d)


def test_config4_like_four_modes():
    # 4 well separated Gaussians in 6-D: the sampler must find 4 ellipsoids /
    # 4 neural bounds and integrate to log Z = 0
    d = 6
    mus = np.full((4, d), 0.5)
    mus[:, 0] = [0.25, 0.25, 0.75, 0.75]
    mus[:, 1] = [0.25, 0.75, 0.25, 0.75]
    like = likelihoods.GaussianMixture(mus, sigma=0.03)
    sampler = Sampler(lambda x: x, like, n_dim=d, n_live=1500, seed=0,
                      n_batch=500)
    assert sampler.run(n_eff=5000, discard_exploration=True)
    last = sampler.bounds[-1]
    # the live set splits into 4 non-overlapping ellipsoids -> 4 neural
    # bounds x 4 networks (nautilus.py:105-114); the outer sampling union is
    # only refined while its volume is far above the target (:123-126)
    assert len(last.neural_bounds) == 4 and last.n_net == 16, (
        len(last.neural_bounds), last.n_net, last.n_ell)
    assert abs(sampler.log_z - like.log_z_true) < 0.1, sampler.log_z
    pts, log_w, _ = sampler.posterior()
    w = np.exp(log_w)
    # each mode carries a quarter of the posterior mass
    for mu in mus:
        near = np.linalg.norm(pts[:, :2] - mu[:2], axis=1) < 0.2
        assert abs(np.sum(w[near]) - 0.25) < 0.05


def test_wide_input_single_group_mode():
    # d = 36 needs more than 256 TMEM columns per tile -> one tile group per
    # CTA; scores must still match the fp64 path and the fused cycle must be
    # consistent with contains()
    d = 36
    rng = np.random.default_rng(0)
    pts = _ball_points(rng, 5000, d, 0.4)
    log_l = -0.5 * np.sum((pts - 0.5)**2, axis=1) / 0.1**2
    log_l_min = np.sort(log_l)[-2000]
    nbound = bounds.NautilusBound.compute(
        pts, log_l, log_l_min, np.log(0.4) * d, n_networks=4,
        rng=np.random.default_rng(1), mode=ops.MLP_TF32)
    spec = nbound.spec()
    meta, _ = pack_stack([spec])
    rec = meta[meta[1]:]
    hdr = rec[rec[rec[8] + 11]:]
    assert hdr[0] >> 16 == 1                       # single tile group
    stack = ops.DeviceStack([spec])
    raw, _, _ = stack.propose(0, 20000, seed=3)
    t = bounds.Ellipsoid.from_matrices(
        spec['neural'][0]['ell']['c'], np.eye(d), np.eye(d))  # placeholder
    del t
    whitened = torch.from_numpy(orc.ell_transform(
        spec['neural'][0]['ell'], raw.cpu().numpy())).cuda()
    p64 = stack.mlp_predict(0, 0, whitened, mode=ops.MLP_F64).cpu().numpy()
    p32 = stack.mlp_predict(0, 0, whitened, mode=ops.MLP_TF32).cpu().numpy()
    assert np.max(np.abs(p64 - p32)) < 5e-3
    like = likelihoods.Gaussian(d)
    out = stack.cycle(0, 1 << 15, seed=5, like_id=like.like_id,
                      like_params=like.device_params('cuda'),
                      mode=ops.MLP_TF32)
    keep, _, n_keep = stack.compact(out['points'], out['log_l'], out['code'])
    n_keep = int(n_keep.item())
    assert n_keep > 0
    assert bool(stack.contains(0, keep[:n_keep].contiguous(),
                               mode=ops.MLP_TF32).all())


def test_config5_like_fixed_bound_100d():
    # 100-D equicorrelated Gaussian, fixed bound built from seeded points; the
    # fp64 cycle must agree with the oracle bit-for-bit at d = 100
    d = 100
    like = likelihoods.EquicorrelatedGaussian(d, sigma=0.05, rho=0.5)
    rng = np.random.default_rng(0)
    cov = like.sigma**2 * ((1 - like.rho) * np.eye(d) + like.rho)
    pts = 0.5 + rng.multivariate_normal(np.zeros(d), cov, size=2500) * 1.2
    pts = pts[np.all((pts > 0) & (pts < 1), axis=1)]
    log_l = like(pts)
    log_l_min = np.sort(log_l)[-800]
    nbound = bounds.NautilusBound.compute(
        pts, log_l, log_l_min, -200.0, n_networks=2,
        rng=np.random.default_rng(1))
    spec = nbound.spec()
    stack = ops.DeviceStack([spec])
    # two default networks at d = 100 still fit shared memory (one tile group,
    # staged fp64 front end); four do not and take the layer-at-a-time
    # tensor-core kernel.  Both must track the fp64 scores.
    from nautilus_b200._pack import pack_tc
    emu2 = spec['neural'][0]['emulator']
    emu4 = dict(emu2, coefs=emu2['coefs'] * 2, intercepts=emu2['intercepts'] * 2)
    assert pack_tc(emu2, 0.5)[0][0] >> 16 == 1
    assert pack_tc(emu4, 0.5)[0][0] >> 16 == 0
    spec4 = dict(spec, neural=[dict(spec['neural'][0], emulator=emu4)])
    raw, _, _ = stack.propose(0, 4096, seed=9)
    whitened = torch.from_numpy(orc.ell_transform(
        spec['neural'][0]['ell'], raw.cpu().numpy())).cuda()
    for st in (stack, ops.DeviceStack([spec4])):
        p64 = st.mlp_predict(0, 0, whitened, mode=ops.MLP_F64).cpu().numpy()
        p32 = st.mlp_predict(0, 0, whitened, mode=ops.MLP_TF32).cpu().numpy()
        assert np.max(np.abs(p64 - p32)) < 5e-3
    n = 4096
    out = stack.cycle(0, n, seed=2, like_id=like.like_id,
                      like_params=like.device_params('cuda'),
                      log_l_min=float(log_l_min), mode=ops.MLP_F64)
    p = out['points'].cpu().numpy()
    code = out['code'].cpu().numpy()
    _, r, _ = orc.replay_integer_stream(n, 0, 0, 2, spec)
    ref_code, _, ref_ll = orc.classify(spec, [], p, r, like)
    assert np.array_equal(code, ref_code)
    sel = code == ops.CODE_IN_SHELL
    assert sel.sum() > 0
    got = out['log_l'].cpu().numpy()[sel]
    assert np.max(np.abs(got - ref_ll[sel]) / np.abs(ref_ll[sel])) < 1e-12


def test_config3_like_wide_network_streamed():
    # config 3's emulator: 50 -> 4 x 128 -> 1, four networks (1 MB of tf32
    # weights): layer-at-a-time tensor-core path against the fp64 kernel
    d, hidden, n_net = 50, (128, 128, 128, 128), 4
    rng = np.random.default_rng(0)
    sizes = (d, ) + hidden + (1, )
    coefs = [[rng.normal(size=(a, b)) * np.sqrt(2.0 / a)
              for a, b in zip(sizes[:-1], sizes[1:])] for _ in range(n_net)]
    intercepts = [[rng.normal(size=b) * 0.1 for b in sizes[1:]]
                  for _ in range(n_net)]
    for n in range(n_net):                  # keep the score O(1)
        coefs[n][-1] *= 0.1
    ell = dict(c=np.full(d, 0.5), B=np.eye(d) * 0.45, B_inv=np.eye(d) / 0.45)
    emu = dict(mean=np.zeros(d), scale=np.full(d, 0.3), coefs=coefs,
               intercepts=intercepts)
    spec = dict(kind='nautilus', n_dim=d, unit=True,
                log_v_all=np.array([0.0]),
                mixtures=[dict(dim_cube=np.zeros(d, bool), ell=ell)],
                neural=[dict(ell=ell, emulator=emu, score_predict_min=0.05)])
    stack = ops.DeviceStack([spec])
    t = torch.from_numpy(rng.normal(size=(5000, d)) * 0.3).cuda()
    p64 = stack.mlp_predict(0, 0, t, mode=ops.MLP_F64).cpu().numpy()
    p32 = stack.mlp_predict(0, 0, t, mode=ops.MLP_TF32).cpu().numpy()
    ref = orc.emulator_predict(emu, t.cpu().numpy())
    assert np.max(np.abs(p64 - ref)) < 1e-12
    err = np.max(np.abs(p32 - p64))
    print('streamed tf32 emulator (4 x 128): max |d score| = {:.2e} at score '
          'scale {:.2f}'.format(err, np.std(p64)))
    assert err < 5e-3 * max(1.0, np.std(p64))
    # the cycle works end to end with the streamed emulator and agrees with
    # contains() under the same arithmetic
    like = likelihoods.Rosenbrock(d)
    out = stack.cycle(0, 1 << 14, seed=1, like_id=like.like_id,
                      like_params=like.device_params('cuda'),
                      mode=ops.MLP_TF32)
    keep, _, n_keep = stack.compact(out['points'], out['log_l'], out['code'])
    n_keep = int(n_keep.item())
    cnt = out['counters'].cpu().numpy()
    assert cnt[ops.CNT_RAW] == 1 << 14 and n_keep == cnt[ops.CNT_IN_SHELL]
    if n_keep:
        assert bool(stack.contains(0, keep[:n_keep].contiguous(),
                                   mode=ops.MLP_TF32).all())


def test_config3_like_wide_network_training():
    # the 4 x 128 emulator of config 3 does not fit shared memory: the trainer
    # keeps weights / gradient in L2-resident global memory ("big" mode)
    from nautilus_b200 import neural
    rng = np.random.default_rng(0)
    d, m = 50, 3000
    x = rng.random((m, d))
    y = np.linalg.norm(x[:, :5] - 0.5, axis=1)
    y = np.argsort(np.argsort(y)) / float(m)
    emu = neural.NeuralNetworkEmulator.train(
        x, y, n_networks=2,
        neural_network_kwargs=dict(hidden_layer_sizes=(128, 128, 128, 128)))
    assert emu.neural_networks[0].coefs_[1].shape == (128, 128)
    pred = emu.predict(x)
    rmse = np.sqrt(np.mean((y - pred)**2))
    print('4 x 128 emulator: rmse / std = {:.3f}, epochs = {}'.format(
        rmse / np.std(y), [n.n_iter_ for n in emu.neural_networks]))
    assert rmse < 0.3 * np.std(y)
    p32 = emu.predict(x, mode=ops.MLP_TF32)
    assert np.max(np.abs(p32 - pred)) < 5e-3


def test_config2_delta_log_z_and_posterior_moments():
    """BASELINE config 2 end to end through the drop-in Sampler: 30-D
    N(0.5, 0.1^2 I), n_live = 2000, tensor-core emulator, device cycle.  Run to
    N_eff >= 1e5 (statistical error 1/sqrt(N_eff) = 0.0032: SURVEY.md 8d asks
    for >= 4e4, where the 0.01 bar is only two sigma) and
    assert north_star's |delta log Z| <= 0.01 against the analytic evidence,
    plus the posterior mean and covariance (the reference's own tolerances,
    tests/test_sampler.py:167-215: 0.01 on the mean, 0.001 on the
    covariance)."""
    d = 30
    like = likelihoods.Gaussian(d, sigma=0.1)
    sampler = Sampler(lambda x: x, like, n_dim=d, n_live=2000, seed=0)
    assert sampler.device_cycle and sampler.mlp_mode == ops.MLP_F16
    assert sampler.run(n_eff=100000, discard_exploration=True, timeout=600)
    assert sampler.n_eff >= 100000
    delta = abs(sampler.log_z - like.log_z_true)
    print('config 2: log Z = {:+.5f} (truth {:+.1e}), |delta| = {:.5f}, '
          'N_eff = {:.0f}, {} bounds, {} likelihood calls'.format(
              sampler.log_z, like.log_z_true, delta, sampler.n_eff,
              len(sampler.bounds), sampler.n_like))
    assert delta <= 0.01
    pts, log_w, _ = sampler.posterior()
    w = np.exp(log_w)
    assert np.isclose(np.sum(w), 1)
    assert np.allclose(np.average(pts, weights=w, axis=0), 0.5, atol=0.01)
    cov = np.cov(pts, aweights=w, rowvar=False)
    assert np.allclose(cov, np.eye(d) * 0.01, atol=0.001)
    # the batch loop read 96 bytes per raw batch from the device
    assert sampler.cycle_stats['d2h_bytes'] == 96 * sampler.cycle_stats['calls']
