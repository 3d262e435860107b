"""End-to-end behaviour of the drop-in Sampler (tests/test_sampler.py of the
reference, re-stated; tolerances are the reference's own)."""

import numpy as np
import pytest
from scipy.stats import multivariate_normal

torch = pytest.importorskip('torch')

from nautilus_b200 import Prior, Sampler, likelihoods  # noqa: E402

pytestmark = pytest.mark.gpu


def test_sampler_accuracy():
    # tests/test_sampler.py:167-215: 2-D Gaussian, sigma 0.1
    n_dim = 2
    mean, cov = np.repeat(0.5, n_dim), np.eye(n_dim) * 0.01

    def likelihood(x):
        return multivariate_normal.logpdf(x, mean=mean, cov=cov)

    sampler = Sampler(lambda x: x, likelihood, n_dim=n_dim, n_live=500,
                      vectorized=True, seed=0)
    assert sampler.run(f_live=0.45, n_eff=0, verbose=False)
    assert sampler.run(n_eff=10000, verbose=False)
    assert np.abs(sampler.log_z) < 0.05
    points, log_w, log_l = sampler.posterior()
    w = np.exp(log_w)
    assert np.isclose(np.sum(w), 1)
    assert np.allclose(np.average(points, weights=w, axis=0), mean, atol=0.01)
    assert np.allclose(np.cov(points, aweights=w, rowvar=False), cov,
                       atol=0.001)
    points, log_w, log_l = sampler.posterior(equal_weight=True)
    assert np.all(log_w == log_w[0]) and 0 < len(points) <= sampler.n_like
    # strictly nested bounds
    occ = sampler.shell_bound_occupation()
    for i in range(len(occ)):
        assert np.all(occ[i, :i + 1] == 1)
        assert np.all(occ[i, i + 1:] < 1)
    assert sampler.n_eff >= 10000
    assert 0 < sampler.eta <= 1


def test_sampler_device_likelihood_matches_host_likelihood():
    # device_cycle=False: the device likelihood evaluated op by op with the
    # step semantics of a host likelihood (exactly n_batch evaluations)
    like = likelihoods.Gaussian(3, mu=[0.4, 0.5, 0.6], sigma=0.1)
    s_dev = Sampler(lambda x: x, like, n_dim=3, n_live=500, seed=1,
                    device_cycle=False)
    s_dev.run(n_eff=3000)
    assert abs(s_dev.log_z - like.log_z_true) < 0.05
    s_host = Sampler(lambda x: x, like.__call__, n_dim=3, n_live=500,
                     vectorized=True, seed=1)
    s_host.run(n_eff=3000)
    # same seed, same arithmetic for log L up to summation order
    assert abs(s_host.log_z - s_dev.log_z) < 1e-9
    assert s_host.n_like == s_dev.n_like


class _StagedCycleSampler(Sampler):
    """The device cycle rebuilt from separate operations: raw draws, neural
    filter, one contains() per later bound, likelihood, and the ORACLE's
    reductions on the host -- everything nb200_cycle fuses into one call."""

    def _run_cycle(self, index, n_raw, offset, out_points, out_log_l):
        from nautilus_b200 import ops
        from oracle import nautilus_oracle as orc
        stack = self._device_stack()
        bound = self.bounds[index]
        pts, code, _ = stack.propose(index, n_raw, seed=bound.stream.seed,
                                     offset=offset,
                                     stream_id=bound.stream.stream_id)
        code = code.clone()
        alive = code == ops.CODE_IN_SHELL
        if index > 0 and len(bound.neural_bounds) > 0:
            nn = stack.contains(index, pts, which=2, mask=alive,
                                mode=self.mlp_mode)
            code[alive & ~nn] = ops.CODE_NN_REJECT
            alive = alive & nn
        for later in range(index + 1, len(self.bounds)):
            # every later bound is consulted (no short-circuit)
            inside = stack.contains(later, pts, mask=alive,
                                    mode=self.mlp_mode)
            code[alive & inside] = ops.CODE_EXCLUDED
        sel = code == ops.CODE_IN_SHELL
        par = self.likelihood.device_params(pts.device)
        log_l = ops.loglike(pts, self.likelihood.like_id, par, code=code)
        k = int(sel.sum().item())
        out_points[:k] = pts[sel]
        out_log_l[:k] = log_l[sel]
        ll = log_l[sel].cpu().numpy()
        cnt = np.zeros(ops.N_CNT, dtype=np.int64)
        cnt[ops.CNT_RAW] = n_raw
        c = code.cpu().numpy()
        for q in range(4):
            cnt[1 + q] = np.sum(c == q)
        cnt[ops.CNT_IN_SHELL] = k
        cnt[ops.CNT_UPDATE] = np.sum(ll >= self.shell_log_l_min[index])
        m, s1, s2 = orc.lse_triple(ll)
        return (torch.as_tensor(cnt, device=pts.device),
                torch.as_tensor(np.array([m, s1, s2, 0.0]),
                                device=pts.device))


def test_device_cycle_equals_staged_operations():
    """add_samples through ONE nb200_cycle call per raw batch (later-bound
    exclusion, likelihood and sums fused) reproduces, shell by shell, the
    run assembled from separate operations and host reductions: identical
    counters and points, log-sum-exp triples to 1e-12."""
    like = likelihoods.Gaussian(4, mu=[0.4, 0.5, 0.6, 0.5], sigma=0.08)
    kw = dict(n_dim=4, n_live=400, seed=3, n_batch=300, emulator_arith='f64')
    a = Sampler(lambda x: x, like, **kw)
    b = _StagedCycleSampler(lambda x: x, like, **kw)
    for s in (a, b):
        assert s.device_cycle
        assert s.run(n_eff=4000, discard_exploration=True)
    assert len(a.bounds) == len(b.bounds) > 3
    assert a.n_like == b.n_like
    assert np.array_equal(a.shell_n, b.shell_n)
    assert np.array_equal(a.shell_n_sample, b.shell_n_sample)
    assert np.allclose(a.shell_log_l, b.shell_log_l, rtol=0, atol=1e-11)
    assert np.allclose(a.shell_n_eff, b.shell_n_eff, rtol=1e-11)
    assert np.allclose(a.shell_log_v, b.shell_log_v, rtol=0, atol=1e-12)
    assert abs(a.log_z - b.log_z) < 1e-11
    for pa, pb in zip(a.points, b.points):
        assert np.array_equal(pa, pb)
    # later-bound exclusion happened inside the cycle, and the host read 96
    # bytes per raw batch
    assert a.cycle_stats['d2h_bytes'] == 96 * a.cycle_stats['calls']
    assert abs(a.log_z - like.log_z_true) < 0.05
    # read-outs from the arena
    pts, log_w, log_l = a.posterior()
    assert len(pts) == np.sum(a.shell_n) and np.isclose(
        np.sum(np.exp(log_w)), 1)
    assert np.allclose(np.average(pts, weights=np.exp(log_w), axis=0),
                       like.mu, atol=0.01)
    assert [len(p) for p in a.points] == list(
        a.shell_n + a.shell_end_exp)


def test_sampler_enlarge_per_dim():
    # tests/test_sampler.py:218-241: bounds so enlarged that every new bound
    # equals the cube and is rejected; analytic log Z of a flat likelihood
    def likelihood(x):
        return -np.linalg.norm(x - 0.5)**2 * 0.001

    sampler = Sampler(lambda x: x, likelihood, n_dim=2, enlarge_per_dim=100,
                      n_networks=0, seed=0)
    sampler.run(f_live=0.1, n_eff=0)
    assert np.isclose(sampler.n_like, sampler.n_eff, rtol=0, atol=1)
    assert len(sampler.bounds) == 1
    assert np.isclose(sampler.log_z, -4 * 0.5**3 / 3 * 0.001, rtol=0,
                      atol=1e-4)


def test_sampler_empty_shells():
    # tests/test_sampler.py:244-258
    def likelihood(x):
        return -np.linalg.norm(x - 0.5)**2 * 0.001

    sampler = Sampler(lambda x: x, likelihood, n_dim=2, n_networks=0, seed=0,
                      n_update=1, n_live=10, n_batch=1)
    sampler.run(f_live=1e-3, n_eff=0)
    assert np.all(sampler.shell_n > 0)


def test_sampler_funnel():
    # tests/test_sampler.py:302-331
    from scipy.stats import norm

    def likelihood(x):
        return (norm.logpdf(x[0], loc=0.5, scale=0.1) +
                norm.logpdf(x[1], loc=0.5, scale=np.exp(20 * (x[0] - 0.5)) /
                            100))

    np.random.seed(0)
    x_0 = np.random.normal(loc=0.5, scale=0.1, size=1000000)
    x_1 = np.random.normal(loc=0.5, scale=np.exp(20 * (x_0 - 0.5)) / 100)
    log_z_true = np.log(np.mean((x_0 > 0) & (x_0 < 1) & (x_1 > 0) & (x_1 < 1)))
    sampler = Sampler(lambda x: x, likelihood, n_dim=2, n_networks=1, seed=0)
    sampler.run()
    assert np.isclose(log_z_true, sampler.log_z, rtol=0, atol=0.1)


def test_sampler_plateau():
    # tests/test_sampler.py:351-369
    def likelihood(x):
        if x[0] < 0.9:
            return -np.inf
        return np.log(x[0] - 0.9)

    log_z_true = np.log(0.5 * 0.1**2)
    for seed in range(3):
        sampler = Sampler(lambda x: x, likelihood, 2, n_live=1000,
                          n_networks=1, seed=seed)
        sampler.run(f_live=0.1)
        assert np.isclose(sampler.log_z, log_z_true, rtol=0, atol=0.1)


def test_sampler_constant_likelihood():
    # tests/test_sampler.py:334-348
    sampler = Sampler(lambda x: x, lambda x: 0, 2, n_live=500, seed=0)
    sampler.run(f_live=0.1, n_eff=0)
    assert np.isclose(sampler.log_z, 0)
    assert len(sampler.bounds) == 1


def test_sampler_n_like_max_and_resume():
    like = likelihoods.Gaussian(2, sigma=0.1)

    def make():
        return Sampler(lambda x: x, like, n_dim=2, n_live=500, seed=0)

    a = make()
    assert a.run(n_eff=3000)
    b = make()
    assert not b.run(n_like_max=700, n_eff=3000)
    assert b.n_like >= 700
    assert b.run(n_eff=3000)
    assert a.log_z == b.log_z and a.n_eff == b.n_eff


def test_sampler_prior_object_and_dict():
    prior = Prior()
    prior.add_parameter('a', (-1, 1))
    prior.add_parameter('b', 0.3)
    prior.add_parameter('c', (0, 2))

    def likelihood(p):
        return -0.5 * ((p['a'] / 0.2)**2 + ((p['c'] - 1) / 0.2)**2) + \
            0 * p['b']

    sampler = Sampler(prior, likelihood, n_live=400, vectorized=True, seed=2)
    sampler.run(n_eff=2000)
    pts, log_w, log_l = sampler.posterior(return_as_dict=True)
    assert set(pts) == {'a', 'b', 'c'} and np.all(pts['b'] == 0.3)
    truth = np.log(2 * np.pi * 0.04 / 4)
    assert abs(sampler.log_z - truth) < 0.1


def test_sampler_errors():
    with pytest.raises(ValueError):
        Sampler(lambda x: x, lambda x: 0.0)                  # n_dim missing
    with pytest.raises(ValueError):
        Sampler(lambda x: x, lambda x: 0.0, n_dim=1)
    with pytest.raises(ValueError):                          # file ending
        Sampler(lambda x: x, lambda x: 0.0, n_dim=2, filepath='x.txt')
    s = Sampler(lambda x: x, lambda x: 0.0, n_dim=2)
    with pytest.raises(ValueError):
        s.discard_exploration = 1


def test_torch_likelihood_plugin_and_device_prior():
    """f-3: a user likelihood written with torch on CUDA tensors + the Prior's
    device transforms keep the whole batch loop on the GPU; same evidence as
    the host-callable run of the same model and as the analytic value."""
    from scipy import stats
    prior = Prior()
    prior.add_parameter('a', (-1.0, 1.0))
    prior.add_parameter('b', stats.norm(loc=0.0, scale=2.0))
    prior.add_parameter('c', (0.0, 2.0))
    calls = {'n': 0}

    def like_torch(p):
        assert p['a'].is_cuda and p['a'].dtype == torch.float64
        calls['n'] += 1
        return -0.5 * ((p['a'] / 0.2)**2 + (p['b'] / 0.3)**2 +
                       ((p['c'] - 1) / 0.2)**2)

    def like_numpy(p):
        return -0.5 * ((p['a'] / 0.2)**2 + (p['b'] / 0.3)**2 +
                       ((p['c'] - 1) / 0.2)**2)

    # Z = int L pi: uniform(-1,1) x N(0,2^2) x uniform(0,2)
    truth = (np.log(np.sqrt(2 * np.pi) * 0.2 / 2) * 2 +
             np.log(0.3 / np.sqrt(0.3**2 + 2.0**2)))
    dev = Sampler(prior, likelihoods.TorchLikelihood(like_torch), n_live=500,
                  seed=4)
    assert dev.device_cycle
    assert dev.run(n_eff=5000, discard_exploration=True)
    assert calls['n'] > 0
    assert abs(dev.log_z - truth) < 0.05, (dev.log_z, truth)
    host = Sampler(prior, like_numpy, n_live=500, vectorized=True, seed=4)
    assert host.run(n_eff=5000, discard_exploration=True)
    assert abs(host.log_z - dev.log_z) < 0.08
    pts, log_w, _ = dev.posterior(return_as_dict=True)
    w = np.exp(log_w)
    assert abs(np.average(pts['c'], weights=w) - 1) < 0.02
    assert abs(np.average(pts['b'], weights=w)) < 0.03
    # plain callable prior + array likelihood
    plain = Sampler(lambda x: x, likelihoods.TorchLikelihood(
        lambda x: -0.5 * torch.sum(((x - 0.5) / 0.1)**2, dim=1)), n_dim=3,
        n_live=400, seed=1)
    assert plain.run(n_eff=3000)
    assert abs(plain.log_z - 3 * np.log(np.sqrt(2 * np.pi) * 0.1)) < 0.06
