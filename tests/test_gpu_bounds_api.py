"""The bound classes behave like the reference's (tests/test_bounds.py and
tests/test_neural.py of johannesulf/nautilus, re-stated for the GPU mirror)."""

import numpy as np
import pytest
from scipy.special import gamma

torch = pytest.importorskip('torch')

from nautilus_b200 import bounds, neural  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture
def points_on_hypersphere_boundary():
    n_dim = 10
    points = np.zeros((2 * n_dim, n_dim)) + 0.5
    for i in range(n_dim * 2):
        points[i, i // 2] += 1 if i % 2 else -1
    return points


@pytest.fixture
def random_points_from_hypersphere():
    np.random.seed(0)
    n_dim, n_points = 3, 1000
    points = np.random.normal(size=(n_points, n_dim))
    points = points / np.sqrt(np.sum(points**2, axis=1))[:, np.newaxis]
    points *= np.random.uniform(size=n_points)[:, np.newaxis]**(1.0 / n_dim)
    return points


@pytest.fixture
def random_points_from_hypercube():
    np.random.seed(0)
    return np.random.random(size=(500, 4))


def test_unit_cube():
    cube = bounds.UnitCube.compute(3)
    points = cube.sample(200)
    assert points.shape == (200, 3)
    assert np.all((points >= 0) & (points < 1))
    assert np.all(cube.contains(points))
    assert cube.log_v == 0


def test_unit_cube_rng():
    n_dim, n_points = 7, 1000
    cube = bounds.UnitCube.compute(n_dim, rng=np.random.default_rng(0))
    same = bounds.UnitCube.compute(n_dim, rng=np.random.default_rng(0))
    diff = bounds.UnitCube.compute(n_dim, rng=np.random.default_rng(1))
    points = cube.sample(n_points)
    assert np.all(points == same.sample(n_points))
    assert not np.all(points == diff.sample(n_points))
    assert not np.all(points == cube.sample(n_points))
    assert np.all(cube.contains(np.random.random((n_points, n_dim))))


def test_ellipsoid_construction():
    with pytest.raises(ValueError):
        bounds.Ellipsoid.compute(np.random.random(size=(10, 10)))
    with pytest.raises(ValueError):
        bounds.Ellipsoid.compute(np.random.random(size=(100, 10)),
                                 enlarge_per_dim=0.9)


def test_mvee_known_answer(points_on_hypersphere_boundary):
    # tests/test_bounds.py:88-101
    from nautilus_b200.bounds._construct import enclosing_ellipsoid
    c_true = np.median(points_on_hypersphere_boundary, axis=0)
    np.random.seed(0)
    points = np.concatenate([points_on_hypersphere_boundary, np.atleast_2d(
        c_true + np.random.random() - 0.5)])
    c, A = enclosing_ellipsoid(points)[:2]
    assert np.allclose(c, c_true, rtol=0, atol=1e-3)
    assert np.allclose(A, np.eye(len(c_true)), rtol=0, atol=1e-2)


def test_ellipsoid_sample_and_contains(points_on_hypersphere_boundary):
    ell = bounds.Ellipsoid.compute(points_on_hypersphere_boundary,
                                   enlarge_per_dim=1.0,
                                   rng=np.random.default_rng(0))
    c = np.mean(points_on_hypersphere_boundary)
    points = ell.sample(100)
    assert points.shape == (100, 10)
    assert np.all(np.linalg.norm(points - c, axis=1) < 1 + 1e-9)
    assert np.all(ell.contains(points))
    ell = bounds.Ellipsoid.compute(points, enlarge_per_dim=1.1,
                                   rng=np.random.default_rng(0))
    points = ell.sample(100)
    assert np.all(ell.contains(points))


def test_ellipsoid_volume(points_on_hypersphere_boundary):
    n_dim = points_on_hypersphere_boundary.shape[1]
    for f in [1.0, 1.1, np.pi / 2.0]:
        ell = bounds.Ellipsoid.compute(points_on_hypersphere_boundary,
                                       enlarge_per_dim=f)
        assert np.isclose(ell.log_v, np.log(
            f**n_dim * np.pi**(n_dim / 2) / gamma(n_dim / 2 + 1)), atol=1e-3)


def test_ellipsoid_transform(random_points_from_hypersphere):
    ell = bounds.Ellipsoid.compute(random_points_from_hypersphere,
                                   rng=np.random.default_rng(0))
    points = ell.sample(100)
    points_t = ell.transform(points)
    assert np.all(np.abs(points_t) < 1 + 1e-9)
    assert np.allclose(points, ell.transform(points_t, inverse=True))
    # single point in, scalar out
    assert ell.contains(points[0]) in (True, np.True_)


def test_ellipsoid_rng(random_points_from_hypersphere):
    mk = lambda s: bounds.Ellipsoid.compute(  # noqa: E731
        random_points_from_hypersphere, rng=np.random.default_rng(s))
    ell, same, diff = mk(0), mk(0), mk(1)
    points = ell.sample(1000)
    assert np.all(points == same.sample(1000))
    assert not np.all(points == diff.sample(1000))
    assert not np.all(points == ell.sample(1000))


def test_mixture_uses_cube_dimensions():
    rng = np.random.default_rng(7)
    pts = rng.random((600, 6))
    pts[:, [1, 3, 4]] = 0.5 + 0.02 * rng.normal(size=(600, 3))
    mix = bounds.UnitCubeEllipsoidMixture.compute(
        pts, rng=np.random.default_rng(0))
    assert list(np.flatnonzero(~mix.dim_cube)) == [1, 3, 4]
    assert np.all(mix.contains(pts))
    smp = mix.sample(500)
    assert np.all(mix.contains(smp))
    assert np.all((smp[:, mix.dim_cube] >= 0) & (smp[:, mix.dim_cube] < 1))
    t = mix.transform(smp)
    assert np.all(np.abs(t[:, mix.dim_cube]) <= 1)
    assert np.all(np.sum(t[:, ~mix.dim_cube]**2, axis=1) < 1)


def test_union_construction():
    with pytest.raises(ValueError):
        bounds.Union.compute(np.random.random(size=(100, 10)),
                             n_points_min=5)


def test_union_split(random_points_from_hypersphere):
    points = np.concatenate([random_points_from_hypersphere,
                             random_points_from_hypersphere + 100,
                             random_points_from_hypersphere + 101])
    union = bounds.Union.compute(points, enlarge_per_dim=1.0 + 1e-9,
                                 unit=False, rng=np.random.default_rng(0))
    while union.split(allow_overlap=False):
        pass
    assert len(union.bounds) == 2
    assert np.all(union.contains(points))
    assert union.split()
    assert len(union.bounds) == 3
    assert np.all(union.contains(points))


def test_union_sample_and_contains(random_points_from_hypersphere):
    union = bounds.Union.compute(random_points_from_hypersphere + 50,
                                 enlarge_per_dim=1.0, unit=False,
                                 rng=np.random.default_rng(0))
    for _ in range(4):
        union.split()
    points = union.sample(100)
    assert points.shape == (100, 3)
    assert np.all(union.contains(points))
    # volume from the reject counters vs the analytic unit ball
    assert abs(union.log_v - np.log(4 / 3 * np.pi)) < 0.1


def test_union_rng(random_points_from_hypersphere):
    def mk(seed):
        u = bounds.Union.compute(random_points_from_hypersphere, unit=False,
                                 rng=np.random.default_rng(seed))
        u.split()
        return u
    union, same, diff = mk(0), mk(0), mk(1)
    points = union.sample(100)
    assert np.all(points == same.sample(100))
    assert not np.all(points == diff.sample(100))
    assert not np.all(points == union.sample(100))


def test_neural_network_emulator():
    # tests/test_neural.py:6-15
    np.random.seed(0)
    x = np.random.random((1000, 5))
    y = np.linalg.norm(x - 0.5, axis=1)
    y = np.argsort(np.argsort(y)) / float(len(y))
    emu = neural.NeuralNetworkEmulator.train(x, y, n_networks=1, pool=None)
    rmse = np.sqrt(np.mean((y - emu.predict(x))**2))
    print('emulator rmse / std = {:.3f}, epochs = {}'.format(
        rmse / np.std(y), emu.neural_networks[0].n_iter_))
    assert rmse < 0.3 * np.std(y)
    assert emu.neural_networks[0].n_layers_ == 5
    assert emu.neural_networks[0].coefs_[0].shape == (5, 100)


def test_trainer_quality_against_sklearn_fits(golden):
    """The CUDA trainer against scikit-learn's own fits of the same data
    (tests/golden/emulator_d5_sklearn_fits.npz: MLPRegressor with
    random_state 0..39 through the reference's train_network,
    nautilus/neural.py:10-32).  The fit is statistical -- other random
    numbers, fp32 -- and sklearn's own seeds spread over a factor of 10 in
    the final loss, so the bar is on the distribution over 40 networks each:
    median final loss, median rmse and the 90th percentile of the rmse within
    1.25x of sklearn's."""
    from nautilus_b200.neural import NeuralNetworkEmulator
    g = golden('emulator_d5')
    ref = golden('emulator_d5_sklearn_fits')
    x, y = g['x'], g['y']
    loss, n_iter, rmse = [], [], []
    for rep in range(4):
        emu = NeuralNetworkEmulator.train(x, y, n_networks=10,
                                          seed=1000 + rep)
        xs = (x - emu.mean) / emu.scale
        for net in emu.neural_networks:
            loss.append(net.loss_)
            n_iter.append(net.n_iter_)
            a = xs
            for i, (w, b) in enumerate(zip(net.coefs_, net.intercepts_)):
                a = a @ w + b
                if i + 1 < len(net.coefs_):
                    a = np.maximum(a, 0)
            rmse.append(np.sqrt(np.mean((a[:, 0] - y)**2)))
    loss, n_iter, rmse = np.array(loss), np.array(n_iter), np.array(rmse)
    print('trainer vs sklearn over 40 networks: median loss {:.2e} vs '
          '{:.2e}, median rmse {:.4f} vs {:.4f}, p90 rmse {:.4f} vs {:.4f}, '
          'median epochs {:.0f} vs {:.0f}'.format(
              np.median(loss), np.median(ref['loss']), np.median(rmse),
              np.median(ref['rmse']), np.percentile(rmse, 90),
              np.percentile(ref['rmse'], 90), np.median(n_iter),
              np.median(ref['n_iter'])))
    assert np.median(loss) <= 1.25 * np.median(ref['loss'])
    assert np.median(rmse) <= 1.25 * np.median(ref['rmse'])
    assert np.percentile(rmse, 90) <= 1.25 * np.percentile(ref['rmse'], 90)
    # the test of the reference itself (tests/test_neural.py:15)
    assert np.all(rmse < 0.3 * np.std(y))


def test_neural_bound_contains(random_points_from_hypercube):
    points = random_points_from_hypercube
    log_l = -np.linalg.norm(points - 0.5, axis=1)
    log_l_min = np.median(log_l)
    nbound = bounds.NeuralBound.compute(points, log_l, log_l_min,
                                        n_networks=1)
    points = np.random.random(size=(1000, 4))
    log_l = -np.linalg.norm(points - 0.5, axis=1)
    in_bound = nbound.contains(points)
    assert np.mean(log_l[in_bound] > log_l_min) >= 0.9


def test_nautilus_bound_sample_and_contains(random_points_from_hypercube):
    points = random_points_from_hypercube
    log_l = -np.linalg.norm(points - 0.5, axis=1)
    log_l_min = np.median(log_l)
    nbound = bounds.NautilusBound.compute(points, log_l, log_l_min,
                                          np.log(0.5), n_networks=1)
    points = np.random.random(size=(1000, 4))
    log_l = -np.linalg.norm(points - 0.5, axis=1)
    assert np.mean(log_l[nbound.contains(points)] > log_l_min) >= 0.9
    points = nbound.sample(100)
    log_l = -np.linalg.norm(points - 0.5, axis=1)
    assert np.mean(log_l > log_l_min) >= 0.9
    assert np.all(nbound.contains(points))


def test_nautilus_bound_gaussian_shell():
    radius, width = 0.45, 0.01
    np.random.seed(0)
    points = np.random.random((10000, 2))
    log_l = -((np.linalg.norm(points - 0.5, axis=1) - radius) / width)**2
    log_l_min = -1
    points, log_l = points[log_l > -100], log_l[log_l > -100]
    log_v_target = np.log(2 * np.pi * radius * width * 2)
    nbound = bounds.NautilusBound.compute(
        points, log_l, log_l_min, log_v_target, split_threshold=1,
        n_networks=1, rng=np.random.default_rng(0))
    points = nbound.sample(10000)
    log_l = -((np.linalg.norm(points - 0.5, axis=1) - radius) / width)**2
    assert np.isclose(nbound.log_v, log_v_target, rtol=0, atol=np.log(2))
    assert np.mean(log_l > log_l_min) > 0.5
    assert nbound.n_net == 1


def test_nautilus_bound_two_peaks():
    np.random.seed(0)
    radius = 1e-5
    points = np.vstack([np.random.normal(size=(1000, 2)) * radius + 0.1,
                        np.random.normal(size=(1000, 2)) * radius + 0.9])

    def likelihood(x):
        return -np.minimum(np.linalg.norm(x - 0.1, axis=-1),
                           np.linalg.norm(x - 0.9, axis=-1)) / radius

    log_l = likelihood(points)
    log_l_min = -1
    log_v_target = np.log(2 * np.pi * radius**2)
    nbound = bounds.NautilusBound.compute(
        points, log_l, log_l_min, log_v_target, n_networks=1,
        rng=np.random.default_rng(0))
    points = nbound.sample(10000)
    log_l = likelihood(points)
    assert np.isclose(nbound.log_v, log_v_target, rtol=0, atol=0.1)
    assert np.mean(log_l > log_l_min) > 0.9
    assert nbound.n_net == 2


def test_nautilus_bound_reset_and_sample(random_points_from_hypercube):
    points = random_points_from_hypercube
    log_l = -np.linalg.norm(points - 0.5, axis=1)
    nbound = bounds.NautilusBound.compute(
        points, log_l, np.median(log_l), np.log(0.5), n_networks=1,
        rng=np.random.default_rng(0))
    nbound.reset(np.random.default_rng(0))
    points_1, volume_1 = nbound.sample(10000), nbound.log_v
    nbound.reset(np.random.default_rng(0))
    points_2, volume_2 = nbound.sample(10000), nbound.log_v
    assert np.all(points_1 == points_2)
    assert volume_1 == volume_2


@pytest.mark.parametrize('n_gpus', [1, 2])
def test_nautilus_bound_gpu_pool(random_points_from_hypercube, n_gpus):
    # tests/test_bounds.py:412-441 (pool n_jobs in {1, 2}): sampling through a
    # pool is reproducible -- and here it is also independent of the pool size
    from nautilus_b200.pool import GpuPool
    if torch.cuda.device_count() < n_gpus:
        pytest.skip('needs {} GPUs'.format(n_gpus))
    points = random_points_from_hypercube
    log_l = -np.linalg.norm(points - 0.5, axis=1)
    nbound = bounds.NautilusBound.compute(
        points, log_l, np.median(log_l), np.log(0.5), n_networks=1,
        rng=np.random.default_rng(0))
    pool = GpuPool(n_gpus)
    nbound.reset(np.random.default_rng(0))
    points_1, volume_1 = nbound.sample(10000, pool=pool), nbound.log_v
    nbound.reset(np.random.default_rng(0))
    points_2, volume_2 = nbound.sample(10000, pool=pool), nbound.log_v
    assert np.all(points_1 == points_2) and volume_1 == volume_2
    nbound.reset(np.random.default_rng(0))
    points_3, volume_3 = nbound.sample(10000), nbound.log_v
    assert np.all(points_1 == points_3) and volume_1 == volume_3


def test_device_mvee_matches_host():
    # k_mvee (persistent Khachiyan kernel) against the host restatement and
    # the reference's known answer (tests/test_bounds.py:88-101)
    from nautilus_b200.bounds import _construct
    rng = np.random.default_rng(0)
    for n, d in [(2000, 30), (500, 4), (300, 10), (40, 3)]:
        z = rng.normal(size=(n, d))
        z *= (rng.uniform(size=(n, 1))**(1.0 / d) /
              np.linalg.norm(z, axis=1)[:, None])
        mix = rng.normal(size=(d, d)) * 0.05 + np.eye(d) * 0.1
        x = 0.5 + z @ mix.T
        _construct._MVEE_CACHE.clear()
        c0, a0, ai0 = _construct.enclosing_ellipsoid(x)
        _construct._MVEE_CACHE.clear()
        c1, a1, ai1 = _construct.enclosing_ellipsoid(x, device='cuda')
        quad = np.einsum('ij,jk,ik->i', x - c1, a1, x - c1)
        assert abs(np.max(quad) - 1) < 1e-9          # touches, encloses all
        assert np.max(np.abs(c1 - c0)) < 1e-6
        assert np.max(np.abs(a1 - a0)) / np.max(np.abs(a0)) < 1e-5
        assert np.allclose(a1 @ ai1, np.eye(d), atol=1e-8)
    # 2d points on the axes at +-1: the unit sphere
    d = 10
    x = np.concatenate([np.eye(d), -np.eye(d)])
    _construct._MVEE_CACHE.clear()
    c, a, _ = _construct.enclosing_ellipsoid(x, device='cuda', tol=1e-6,
                                             max_updates=20000)
    assert np.allclose(c, 0, atol=1e-3) and np.allclose(a, np.eye(d),
                                                        atol=1e-2)


@pytest.mark.parametrize('d', [3, 10, 30])
def test_device_mvee_against_the_reference_ellipsoid(golden, d):
    """The device enclosing ellipsoid against the REFERENCE's own
    (tests/golden/ellipsoid_d*.npz stores c, A, log_v of
    Ellipsoid.compute(pts, enlarge_per_dim=1.1) for a seeded point cloud that
    is regenerated here): both enclose every point; the volumes agree to a few
    per cent (Khachiyan's iteration is stopped by tolerance in the reference,
    basic.py:175-241, and run to 3000 updates here: slightly tighter)."""
    g = golden('ellipsoid_d{}'.format(d))
    rng = np.random.default_rng(100 + d)
    L = np.tril(rng.normal(size=(d, d))) * 0.3 + np.eye(d)
    pts = 0.5 + 0.05 * rng.normal(size=(40 * d, d)) @ L.T / np.sqrt(d)
    # the fixture's ellipsoid encloses the regenerated cloud (same cloud)
    ref_quad = np.einsum('ij,jk,ik->i', pts - g['c'], g['A'], pts - g['c'])
    assert np.max(ref_quad) <= 1.0 / 1.1**2 + 1e-9
    ell = bounds.Ellipsoid.compute(pts, enlarge_per_dim=1.1,
                                   rng=np.random.default_rng(0))
    quad = np.einsum('ij,jk,ik->i', pts - ell.c, ell.A, pts - ell.c)
    assert abs(np.max(quad) - 1.0 / 1.1**2) < 1e-8       # touches, encloses
    assert bool(np.all(ell.contains(pts)))
    print('d = {}: log_v device {:.5f} reference {:.5f}'.format(
        d, ell.log_v, float(g['log_v'])))
    assert ell.log_v <= float(g['log_v']) + 1e-3          # never looser
    assert float(g['log_v']) - ell.log_v < 0.03 * d       # ... nor far off
    assert np.max(np.abs(ell.c - g['c'])) < 0.02


def test_tensor_core_trainer_follows_the_simt_trainer(golden, monkeypatch):
    """k_mlp_fit_tc (tcgen05 kind::tf32 products, fp32 accumulate) against
    k_mlp_fit (fp32 FFMA) from the same initial networks through the same
    minibatches: after a few epochs the weights agree to tf32 rounding noise
    and the epoch losses to a few per cent; run to the end both meet the
    reference's own bar."""
    from nautilus_b200.neural import NeuralNetworkEmulator
    g = golden('emulator_d5')
    x, y = g['x'], g['y']

    def fit(kind, **kw):
        monkeypatch.setenv('NB200_FIT', kind)
        return NeuralNetworkEmulator.train(x, y, n_networks=3, seed=7,
                                           neural_network_kwargs=kw)

    for epochs in (1, 4):
        a = fit('tc', max_iter=epochs)
        b = fit('ffma', max_iter=epochs)
        for na, nb_ in zip(a.neural_networks, b.neural_networks):
            assert na.n_iter_ == nb_.n_iter_ == epochs
            assert abs(na.loss_ / nb_.loss_ - 1) < 0.05, (na.loss_, nb_.loss_)
            for wa, wb in zip(na.coefs_ + na.intercepts_,
                              nb_.coefs_ + nb_.intercepts_):
                # (Adam's first steps move every weight by ~lr = 0.01 in the
                # direction of the gradient's SIGN: a weight whose gradient is
                # at rounding level may go the other way, hence the max bar)
                diff = np.abs(wa - wb)
                assert np.mean(diff) < 2e-3 * epochs, (epochs, np.mean(diff))
                if epochs == 1:        # later the trajectories drift apart
                    assert np.max(diff) < 0.1, np.max(diff)
    a, b = fit('tc'), fit('ffma')
    for emu in (a, b):
        rmse = np.sqrt(np.mean((emu.predict(x) - y)**2))
        assert rmse < 0.1 * np.std(y)
    print('final losses tc {} ffma {}'.format(
        [round(n.loss_, 6) for n in a.neural_networks],
        [round(n.loss_, 6) for n in b.neural_networks]))


@pytest.mark.gpu
@pytest.mark.parametrize('hidden', [(32, ), (47, ), (48, ), (64, 16),
                                    (40, 40), (24, 40, 12)])
def test_tensor_core_trainer_other_network_shapes(golden, monkeypatch,
                                                  hidden):
    """The other instantiations of k_mlp_fit_tc (1 and 2 hidden layers), a
    fan-out whose padded width grows by the spare constant-one row (47 -> 48,
    48 -> 64 columns) and a last hidden layer wider than 31 units (the
    4-chunk output stage): one epoch from the same initial networks through
    the same minibatches as the SIMT trainer, then a full fit."""
    from nautilus_b200.neural import NeuralNetworkEmulator
    g = golden('emulator_d5')
    x, y = g['x'], g['y']

    def fit(kind, **kw):
        monkeypatch.setenv('NB200_FIT', kind)
        return NeuralNetworkEmulator.train(
            x, y, n_networks=2, seed=11,
            neural_network_kwargs=dict(hidden_layer_sizes=hidden, **kw))

    a, b = fit('tc', max_iter=1), fit('ffma', max_iter=1)
    largest = 0.0
    for na, nb_ in zip(a.neural_networks, b.neural_networks):
        assert na.n_iter_ == nb_.n_iter_ == 1
        assert abs(na.loss_ / nb_.loss_ - 1) < 0.05, (na.loss_, nb_.loss_)
        for wa, wb in zip(na.coefs_ + na.intercepts_,
                          nb_.coefs_ + nb_.intercepts_):
            assert wa.shape == wb.shape
            diff = np.abs(wa - wb)
            assert np.mean(diff) < 2e-3, np.mean(diff)
            assert np.max(diff) < 0.1, np.max(diff)
            largest = max(largest, float(np.max(diff)))
    # tf32 products leave a trace: bit-identical weights would mean that the
    # shape fell outside the tensor-core envelope and both fits were SIMT
    assert largest > 0.0
    a = fit('tc')
    rmse = np.sqrt(np.mean((a.predict(x) - y)**2))
    assert rmse < 0.25 * np.std(y), rmse
