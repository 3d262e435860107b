"""Host-side logic that needs no GPU: Prior (tests/test_prior.py of the
reference, re-stated), bound serialisation, the tensor-core weight layout and
the construction helpers."""

import numpy as np
import pytest
from scipy.stats import norm

from nautilus_b200._pack import (_core_matrix_layout, flat_to_spec, pack_stack,
                                 pack_tc, spec_to_flat, tf32_round, union_cdf)
from nautilus_b200.bounds import _construct
from nautilus_b200.prior import Prior
from oracle import nautilus_oracle as orc


# ---- Prior ---------------------------------------------------------------

def test_prior_add_parameter_errors():
    with pytest.raises(TypeError):
        Prior().add_parameter(1.0)
    prior = Prior()
    prior.add_parameter('a')
    prior.add_parameter('b')
    with pytest.raises(ValueError):
        prior.add_parameter('a')
    with pytest.raises(ValueError):
        Prior().add_parameter('a', dist='a')
    with pytest.raises(TypeError):
        Prior().add_parameter(dist=[0.0])


def test_prior_dimensionality():
    prior = Prior()
    for _ in range(6):
        prior.add_parameter()
    assert prior.dimensionality() == 6
    assert prior.keys[3] == 'x_3'
    prior = Prior()
    prior.add_parameter(key='a')
    prior.add_parameter(key='b', dist=1)
    prior.add_parameter(key='c', dist='a')
    prior.add_parameter(key='d', dist='b')
    prior.add_parameter(key='e')
    assert prior.dimensionality() == 2


def test_prior_unit_to_physical_and_dictionary():
    prior = Prior()
    prior.add_parameter(key='a', dist=(-1, +1))
    prior.add_parameter(key='b', dist='a')
    prior.add_parameter(key='c', dist='b')
    d1, d2 = norm(loc=3.0, scale=2.0), norm(loc=5.0, scale=1.0)
    prior.add_parameter(key='d', dist=d1)
    prior.add_parameter(key='e', dist=d2)
    prior.add_parameter(key='f', dist=0.5)
    rng = np.random.default_rng(0)
    with pytest.raises(ValueError):
        prior.unit_to_physical(rng.random(4))
    for shape in [(3, ), (10, 3)]:
        unit = rng.random(shape)
        phys = prior.unit_to_physical(unit)
        assert phys.shape == unit.shape
        assert np.allclose(phys[..., 0], unit[..., 0] * 2 - 1)
        assert np.allclose(phys[..., 1], d1.isf(1 - unit[..., 1]))
        assert np.allclose(phys[..., 2], d2.isf(1 - unit[..., 2]))
        out = prior.unit_to_dictionary(unit)
        assert set(out) == {'a', 'b', 'c', 'd', 'e', 'f'}
        assert np.all(out['b'] == out['a']) and np.all(out['c'] == out['a'])
        assert np.all(out['f'] == 0.5)
        assert np.all(out['d'] == phys[..., 1])
    with pytest.raises(ValueError):
        prior.physical_to_dictionary(rng.random(4))


# ---- serialisation ---------------------------------------------------------

def test_spec_flat_round_trip(golden):
    for name in ('union_d5', 'cfg2_bound_d30', 'mixture_d6'):
        g = golden(name)
        spec = flat_to_spec(g)
        again = flat_to_spec(spec_to_flat(spec))
        m1, d1 = pack_stack([spec])
        m2, d2 = pack_stack([again])
        assert np.array_equal(m1, m2) and np.array_equal(d1, d2)


def test_pack_stack_layout(golden):
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    cube = dict(kind='cube', n_dim=30)
    meta, data = pack_stack([cube, spec, spec])
    assert meta[0] == 3
    r0, r1, r2 = (meta[meta[1 + i]:] for i in range(3))
    assert (r0[1], r0[2]) == (0, 30)
    assert (r1[1], r1[2], r1[3], r1[4], r1[5]) == (1, 30, 1, 1, 1)
    assert r1[10] == 1           # neural ellipsoid == mixture 0 (shared pass)
    mix = r1[r1[7]:r1[7] + 8]
    assert (mix[0], mix[1], mix[6]) == (30, 0, 1)
    ell = spec['mixtures'][0]['ell']
    assert np.array_equal(data[mix[3]:mix[3] + 30], ell['c'])
    assert np.array_equal(data[mix[5]:mix[5] + 900], ell['B_inv'].ravel())
    nb = r1[r1[8]:r1[8] + 12]
    assert (nb[3], nb[4]) == (4, 4)
    assert data[nb[7]] == spec['neural'][0]['score_predict_min'] - 1e-9
    # the second copy owns different data offsets
    assert r2[r2[7] + 3] != mix[3]
    assert np.allclose(union_cdf(spec['log_v_all']), [1.0])
    with pytest.raises(ValueError):
        pack_stack([dict(kind='cube', n_dim=129)])


def test_tensor_core_blob_layout(golden):
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    emu = spec['neural'][0]['emulator']
    hdr, blob = pack_tc(emu, 0.5)
    n_groups, n_net, n_hid, d, k0p = hdr[0] >> 16, hdr[1], hdr[2], hdr[3], hdr[4]
    assert (hdr[0] & 0xFFFF, n_groups, n_net, n_hid, d, k0p) == (
        0x7F32, 2, 4, 3, 30, 32)
    np_, kp = hdr[8:11], hdr[12:15]
    assert list(np_) == [112, 64, 32] and list(kp) == [32, 104, 56]
    # element (n, k) of layer l sits at the core-matrix address the kernel's
    # smem descriptor describes: (n/8)*(8*kp) + (k/4)*32 + (n%8)*4 + (k%4)
    w_off, stride = hdr[16:19], hdr[5]
    for net in (0, 3):
        for lay in range(3):
            w = emu['coefs'][net][lay]
            b = emu['intercepts'][net][lay]
            fi, fo = w.shape

            def at(n, k):
                return blob[net * stride + w_off[lay] + (n // 8) * 8 * kp[lay]
                            + (k // 4) * 32 + (n % 8) * 4 + (k % 4)]
            for n, k in [(0, 0), (7, 3), (8, 4), (fo - 1, fi - 1), (13, 17)]:
                assert at(n, k) == tf32_round(np.float32(w[k, n]))
            # bias rides in the constant-one column, which is regenerated
            assert at(5, fi) == tf32_round(np.float32(b[5]))
            assert at(fo, fi) == (1.0 if lay < 2 else 0.0)
            assert at(fo, 0) == 0.0
    thr = np.array(hdr[30:32], dtype=np.int32).view(np.float64)[0]
    assert thr == 0.5 - 1e-9
    lay = _core_matrix_layout(np.arange(16 * 8, dtype=np.float32).reshape(
        16, 8), 16, 8)
    assert lay[0] == 0 and lay[4] == 8 and lay[32] == 4 and lay[64] == 64
    # config 3 (50 -> 4 x 128 -> 1, four networks): 1 MB of weights -> the
    # streamed (layer-at-a-time) mode; five hidden layers -> no blob at all
    def zeros(sizes, n_net=4):
        return dict(coefs=[[np.zeros((a, b)) for a, b in zip(
            sizes[:-1], sizes[1:])]] * n_net,
            intercepts=[[np.zeros(b) for b in sizes[1:]]] * n_net)
    hdr3, blob3 = pack_tc(zeros((50, 128, 128, 128, 128, 1)), 0.0)
    assert hdr3[0] >> 16 == 0 and list(hdr3[8:12]) == [144] * 4
    assert list(hdr3[12:16]) == [56, 136, 136, 136]
    assert pack_tc(zeros((50, 64, 64, 64, 64, 64, 1)), 0.0) is None


def test_tf32_round():
    x = np.array([1.0, 1.0 + 2.0**-11, 1.0 + 2.0**-10, -3.1415927, 0.0],
                 dtype=np.float32)
    r = tf32_round(x)
    assert r[0] == 1.0 and r[2] == x[2] and r[4] == 0.0
    assert r[1] == np.float32(1.0 + 2.0**-10)        # ties away from zero
    assert abs(r[3] - x[3]) <= abs(x[3]) * 2.0**-11
    assert np.all((r.view(np.uint32) & 0x1FFF) == 0)


# ---- construction helpers ----------------------------------------------------

def test_enclosing_ellipsoid_known_answer():
    # tests/test_bounds.py:88-101 of the reference
    n_dim = 10
    points = np.zeros((2 * n_dim, n_dim)) + 0.5
    for i in range(n_dim * 2):
        points[i, i // 2] += 1 if i % 2 else -1
    np.random.seed(0)
    points = np.concatenate([points, np.atleast_2d(
        np.full(n_dim, 0.5) + np.random.random() - 0.5)])
    c, A, A_inv = _construct.enclosing_ellipsoid(points)
    assert np.allclose(c, 0.5, rtol=0, atol=1e-3)
    assert np.allclose(A, np.eye(n_dim), rtol=0, atol=1e-2)
    assert np.allclose(A @ A_inv, np.eye(n_dim), atol=1e-10)
    maha = np.einsum('ij,jk,ik->i', points - c, A, points - c)
    assert np.isclose(maha.max(), 1.0)


def test_enclosing_ellipsoid_elongated_cloud():
    np.random.seed(0)
    radius = 1e-5
    points = np.vstack([np.random.normal(size=(500, 2)) * radius + 0.1,
                        np.random.normal(size=(500, 2)) * radius + 0.9])
    c, A, A_inv = _construct.enclosing_ellipsoid(points)
    maha = np.einsum('ij,jk,ik->i', points - c, A, points - c)
    assert np.all(np.isfinite(A)) and np.isclose(maha.max(), 1.0)
    assert np.allclose(c, 0.5, atol=0.05)


def test_two_gaussians_and_overlap():
    rng = np.random.default_rng(0)
    x = np.vstack([rng.normal(size=(300, 4)), rng.normal(size=(300, 4)) + 8])
    log_p = _construct.two_gaussians(x, rng)
    labels = np.argmax(log_p, axis=1)
    assert len(set(labels[:300])) == 1 and len(set(labels[300:])) == 1
    assert labels[0] != labels[-1]

    class E:
        def __init__(self, c, r):
            self.c = np.asarray(c, dtype=float)
            self.A = np.eye(len(c)) / r**2
    assert not _construct.ellipsoids_overlap([E([0, 0], 1), E([2.5, 0], 1)])
    assert _construct.ellipsoids_overlap([E([0, 0], 1), E([1.5, 0], 1)])
    assert _construct.ellipsoids_overlap(
        [E([0, 0], 1), E([5, 5], 1), E([5.5, 5], 1)])


def test_oracle_classify_is_consistent_with_contains(golden):
    # the disposition codes of oracle.classify agree with bound_contains
    spec = flat_to_spec(golden('nautilus_d4'))
    rng = np.random.default_rng(1)
    pts = rng.random((2000, 4)) * 1.1 - 0.05
    code, nb, _ = orc.classify(spec, [], pts, np.full(len(pts), 0.999))
    inside = orc.bound_contains(spec, pts)
    assert np.array_equal(code == 4, inside)
    assert np.all(nb[code == 0] == 0)


def test_batched_gmm_matches_sequential():
    # the device EM (all restarts advanced together) follows the sequential
    # host EM restart by restart: same seeding draws, same winner, same split
    from nautilus_b200.bounds import _construct
    rng = np.random.default_rng(3)
    for n, d, k in [(900, 6, 3), (400, 3, 2), (300, 10, 1)]:
        centres = rng.uniform(0.2, 0.8, size=(k, d))
        x = np.concatenate([c + 0.03 * rng.normal(size=(n // k, d))
                            for c in centres])
        a = _construct.two_gaussians(x, np.random.default_rng(5))
        b = _construct.two_gaussians_batched(x, np.random.default_rng(5),
                                             device='cpu')
        assert np.max(np.abs(a - b)) < 1e-8
        assert np.array_equal(a.argmax(axis=1), b.argmax(axis=1))
    # degenerate input: fewer points than a covariance needs
    x = rng.normal(size=(4, 6))
    a = _construct.two_gaussians(x, np.random.default_rng(1))
    b = _construct.two_gaussians_batched(x, np.random.default_rng(1),
                                         device='cpu')
    assert np.array_equal(a, b)


def test_prior_device_transforms_match_scipy():
    """Prior.unit_to_*_device (torch) == the SciPy isf path of the reference
    (nautilus/prior.py:85-181); torch CPU tensors exercise the same code the
    CUDA tensors take."""
    import torch
    from scipy import stats
    from nautilus_b200 import Prior
    prior = Prior()
    prior.add_parameter('a', (-1.0, 2.0))
    prior.add_parameter('fixed', 0.25)
    prior.add_parameter('b', stats.norm(loc=2.0, scale=0.5))
    prior.add_parameter('c', stats.expon(scale=3.0))
    prior.add_parameter('d', stats.lognorm(0.4, loc=1.0, scale=2.0))
    prior.add_parameter('tied', 'b')
    u = np.random.default_rng(0).random((257, 4))
    ref = prior.unit_to_physical(u)
    got = prior.unit_to_physical_device(torch.from_numpy(u)).numpy()
    assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref))) < 1e-12
    ref_d = prior.unit_to_dictionary(u)
    got_d = prior.unit_to_dictionary_device(torch.from_numpy(u))
    assert set(ref_d) == set(got_d)
    for key in ref_d:
        assert np.allclose(got_d[key].numpy(), ref_d[key], rtol=1e-12)
    with pytest.raises(ValueError):
        prior.unit_to_physical_device(torch.zeros((3, 5), dtype=torch.float64))
    other = Prior()
    other.add_parameter('x', stats.beta(2, 3))
    with pytest.raises(NotImplementedError):
        other.unit_to_physical_device(torch.zeros((3, 1),
                                                  dtype=torch.float64))


def test_projection_scan_matches_the_per_candidate_form():
    """The dimension search of UnitCubeEllipsoidMixture.compute
    (nautilus/bounds/basic.py:497-512) restated per candidate -- delete the
    column, invert the projected shape matrix, rescale to the farthest point,
    log det -- against the one-product closed form the device path uses."""
    import torch
    rng = np.random.default_rng(3)
    for n, m in ((400, 2), (600, 7), (2000, 30)):
        pts = rng.normal(size=(n, m)) @ rng.normal(size=(m, m))
        c = pts.mean(axis=0) + 0.05 * rng.normal(size=m)
        a = np.linalg.inv(np.cov(pts.T)) / 25.0
        a_inv = np.linalg.inv(a)
        ref = np.zeros(m)
        for i in range(m):
            pts_proj = np.delete(pts, i, axis=1)
            c_proj = np.delete(c, i)
            a_proj = np.linalg.inv(
                np.delete(np.delete(a_inv, i, axis=0), i, axis=1))
            diff = pts_proj - c_proj
            a_proj = a_proj / np.amax(np.einsum('...i,ij,...j', diff, a_proj,
                                                diff))
            ref[i] = np.linalg.slogdet(np.linalg.inv(a_proj))[1]
        got = _construct.projection_scan(
            torch.from_numpy(pts), torch.from_numpy(c),
            torch.from_numpy(a)).numpy()
        assert np.allclose(2 * got, ref, rtol=0, atol=1e-10 * m)
        assert np.argmin(got) == np.argmin(ref)


def test_batched_overlap_test_matches_golden_section():
    """All pairs at once through the generalised eigen-decomposition vs the
    d x d solve per evaluation; also the value of min K itself."""
    import torch

    class Ell:
        pass

    rng = np.random.default_rng(5)
    m = 6
    decided = 0
    for trial in range(120):
        ells = []
        for k in range(3):
            e = Ell()
            mat = rng.normal(size=(m, m))
            e.A = (mat @ mat.T + m * np.eye(m)) / 4.0
            e.c = rng.normal(size=m) * rng.uniform(0.05, 0.8)
            ells.append(e)
        host = _construct._ellipsoids_overlap_host(ells)
        dev = _construct.ellipsoids_overlap(ells, device='cpu')
        assert host == dev
        decided += host
        # min K of one pair against a dense scan of the definition
        a_inv = [np.linalg.inv(e.A) for e in ells[:2]]
        d = ells[0].c - ells[1].c
        s = np.linspace(1e-6, 1 - 1e-6, 20001)
        k_def = np.array([1 - d @ np.linalg.solve(
            a_inv[0] / (1 - x) + a_inv[1] / x, d) for x in s[::50]])
        k_min = _construct.overlap_k_min(
            torch.from_numpy(np.stack([e.c for e in ells])),
            torch.from_numpy(np.stack([np.linalg.inv(e.A) for e in ells])),
            [(0, 1)])[0].item()
        assert k_min <= k_def.min() + 1e-12
        assert k_min >= k_def.min() - 1e-2 * max(1.0, abs(k_def.min()))
    assert 10 < decided < 110          # both outcomes exercised


def test_record_header_same_k_and_rounding_amplification(golden):
    """Record header words 10 / 11 (include/nautilus_b200.h): which mixture's
    ellipsoid is the neural bound's, and ceil(log2(|B_inv|_2 (|B|_2 +
    max|c|))) of it -- the number the front kernel's whitening shortcut
    derives its guard band from."""
    spec = flat_to_spec(golden('cfg2_bound_d30'))
    meta, _ = pack_stack([spec])
    rec = meta[meta[1]:]
    ell = spec['mixtures'][0]['ell']
    amp = np.linalg.norm(ell['B_inv'], 2) * (
        np.linalg.norm(ell['B'], 2) + np.max(np.abs(ell['c'])))
    assert rec[10] == 1 and rec[11] == int(np.ceil(np.log2(amp)))
    # a distinct neural ellipsoid: no shortcut
    spec['neural'][0]['ell']['c'] = spec['neural'][0]['ell']['c'] + 1e-3
    meta, _ = pack_stack([spec])
    rec = meta[meta[1]:]
    assert rec[10] == 0 and rec[11] == 0
    # an ill-conditioned shared ellipsoid: the band would be wide, the
    # launcher turns the shortcut off above 1e-4 (256 d eps 2^amp_log2)
    d = 6
    B = np.diag(np.logspace(0, -12, d)) * 0.3
    ell = dict(c=np.full(d, 0.5), B=B, B_inv=np.diag(1 / np.diag(B)))
    spec = dict(kind='nautilus', n_dim=d, unit=True, log_v_all=np.zeros(1),
                mixtures=[dict(dim_cube=np.zeros(d, bool), ell=ell)],
                neural=[dict(ell={k: v.copy() for k, v in ell.items()},
                             emulator=None, score_predict_min=0.0)])
    meta, _ = pack_stack([spec])
    rec = meta[meta[1]:]
    assert rec[10] == 1 and rec[11] >= 40
    assert 256 * d * 2.0**-53 * 2.0**rec[11] > 1e-4
