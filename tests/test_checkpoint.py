"""Checkpoints (nautilus/sampler.py:329-371,
1253-1377 and the write / update / read methods of every bound).

CPU part: the built-in container and the LAYOUT -- a bound written by the
unmodified reference (oracle/_ref) is read by this package and vice versa; the
reference only touches a group through ``attrs`` / ``create_group`` /
``create_dataset`` / ``[]`` / ``in``, so it writes into the container
unchanged.  GPU part: a run that is stopped, stored, and resumed in a new
``Sampler`` ends where the uninterrupted run ends.
"""

import numpy as np
import pytest

from nautilus_b200 import Prior, Sampler, likelihoods
from nautilus_b200 import _store
from nautilus_b200._pack import flat_to_spec
from nautilus_b200.bounds import NautilusBound


# ---------------------------------------------------------------- container

def test_store_roundtrip(tmp_path):
    path = tmp_path / 'a.npz'
    f = _store.open_store(path, 'x')
    g = f.create_group('sampler')
    g.attrs['n_dim'] = 3
    g.attrs['explored'] = True
    g.attrs['rng_state'] = str(2**127 + 12345)
    g.attrs['shell_log_l'] = np.array([np.nan, -np.inf, 1.5])
    g.attrs['empty'] = np.zeros(0, dtype=int)
    g.create_dataset('points_0', data=np.zeros((0, 3)), maxshape=(None, 3))
    g.create_group('nothing')
    sub = f.create_group('bound_0').create_group('cube')
    sub.attrs['type'] = 'UnitCube'
    rec = np.zeros(4, dtype=[('a', float), ('b', int)])
    rec['b'] = np.arange(4)
    g.create_dataset('blobs_0', data=rec)
    with pytest.raises(ValueError):
        g.create_dataset('points_0', data=np.zeros(1))
    with pytest.raises(TypeError):
        g.attrs['bad'] = dict(a=1)
    f.close()

    with pytest.raises(FileExistsError):
        _store.open_store(path, 'x')
    with _store.open_store(path, 'r') as f:
        g = f['sampler']
        assert g.attrs['n_dim'] == 3 and g.attrs['explored']
        assert int(g.attrs['rng_state']) == 2**127 + 12345
        assert np.array_equal(g.attrs['shell_log_l'],
                              [np.nan, -np.inf, 1.5], equal_nan=True)
        assert g.attrs['empty'].shape == (0, )
        assert 'nothing' in g and 'points_0' in g and 'points_1' not in g
        assert 'sampler/points_0' in f and 'bound_0/cube' in f
        assert f['bound_0']['cube'].attrs['type'] == 'UnitCube'
        assert np.array(g['points_0']).shape == (0, 3)
        assert np.array_equal(np.array(g['blobs_0'])['b'], np.arange(4))

    with _store.open_store(path, 'r+') as f:
        ds = f['sampler']['points_0']
        ds.resize((2, 3))
        ds[...] = np.arange(6.0).reshape(2, 3)
        f['sampler'].attrs['n_like'] = 7
    with _store.open_store(path) as f:
        assert np.array_equal(np.array(f['sampler']['points_0']),
                              np.arange(6.0).reshape(2, 3))
        assert f['sampler'].attrs['n_like'] == 7
    # an interrupted write leaves the previous file alone: nothing but the
    # final rename touches it
    assert not (tmp_path / 'a.npz.tmp').exists()


def test_store_suffixes(tmp_path):
    with pytest.raises(ValueError):
        _store.open_store(tmp_path / 'a.txt', 'w')
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match='h5py'):
            _store.open_store(tmp_path / 'a.hdf5', 'w')
        with pytest.raises(ImportError, match='h5py'):
            _store.require_backend(tmp_path / 'a.h5')
    _store.require_backend(tmp_path / 'a.npz')


# ------------------------------------------------------------------- layout

def _same_spec(a, b):
    assert a['n_dim'] == b['n_dim'] and bool(a['unit']) == bool(b['unit'])
    assert np.array_equal(a['log_v_all'], b['log_v_all'])
    assert len(a['mixtures']) == len(b['mixtures'])
    for ma, mb in zip(a['mixtures'], b['mixtures']):
        assert np.array_equal(ma['dim_cube'], mb['dim_cube'])
        assert (ma['ell'] is None) == (mb['ell'] is None)
        if ma['ell'] is not None:
            for key in ('c', 'B', 'B_inv'):
                assert np.array_equal(ma['ell'][key], mb['ell'][key])
    assert len(a['neural']) == len(b['neural'])
    for na, nb in zip(a['neural'], b['neural']):
        assert na['score_predict_min'] == nb['score_predict_min']
        for key in ('c', 'B', 'B_inv'):
            assert np.array_equal(na['ell'][key], nb['ell'][key])
        ea, eb = na['emulator'], nb['emulator']
        assert np.array_equal(ea['mean'], eb['mean'])
        assert np.array_equal(ea['scale'], eb['scale'])
        for key in ('coefs', 'intercepts'):
            assert len(ea[key]) == len(eb[key])
            for wa, wb in zip(ea[key], eb[key]):
                assert len(wa) == len(wb)
                assert all(np.array_equal(x, y) for x, y in zip(wa, wb))


@pytest.mark.parametrize('name', ['cfg2_bound_d30', 'nautilus_d4'])
def test_bound_layout_is_the_references(golden, tmp_path, name):
    """reference.write -> this package's read -> this package's write ->
    reference.read: parameters identical at every step, and the reference
    bound that comes out of the file answers contains() as the fixture
    recorded."""
    from oracle import ref_arm
    if not ref_arm.available():
        pytest.skip('oracle/_ref missing (run oracle/make_ref.sh)')
    g = golden(name)
    spec = flat_to_spec(g)
    ref_bound = ref_arm.reference_bound(spec, np.random.default_rng(0))
    ref_bound.n_sample, ref_bound.n_reject = 1000, 250
    ref_bound.outer_bound.n_sample = 1234
    ref_bound.outer_bound.n_reject = 234

    path = tmp_path / 'ref.npz'
    with _store.open_store(path, 'x') as f:
        ref_bound.write(f.create_group('bound_1'))       # the reference's
    with _store.open_store(path) as f:
        mine = NautilusBound.read(f['bound_1'], rng=np.random.default_rng(1))
    _same_spec(mine.spec(), spec)
    assert (mine.n_sample, mine.n_reject) == (1000, 250)
    assert (mine.outer_bound.n_sample, mine.outer_bound.n_reject) == \
        (1234, 234)
    assert mine.n_net == ref_bound.n_net and mine.n_ell == ref_bound.n_ell

    path = tmp_path / 'mine.npz'
    mine.stream.take(4096)
    with _store.open_store(path, 'x') as f:
        mine.write(f.create_group('bound_1'))
    with _store.open_store(path) as f:
        import nautilus                           # oracle/_ref (ref_arm)
        back = nautilus.bounds.NautilusBound.read(
            f['bound_1'], rng=np.random.default_rng(2))
        again = NautilusBound.read(f['bound_1'],
                                   rng=np.random.default_rng(3))
    assert np.array_equal(back.contains(g['points']), g['contains'])
    assert (back.n_sample, back.outer_bound.n_reject) == (1000, 234)
    _same_spec(again.spec(), spec)
    # the Philox stream of the bound travels with it
    assert (again.stream.seed, again.stream.stream_id, again.stream.offset) \
        == (mine.stream.seed, mine.stream.stream_id, 4096)


# ------------------------------------------------------------------- resume

@pytest.mark.gpu
def test_resume_continues_the_run(tmp_path):
    like = likelihoods.Gaussian(4, sigma=0.1)
    path = tmp_path / 'run.npz'

    def make(**kwargs):
        return Sampler(lambda x: x, like, n_dim=4, n_live=500, seed=3,
                       **kwargs)

    a = make()
    assert a.run(n_eff=4000, discard_exploration=True)

    # stopped during the exploration phase, stored, resumed in a new object
    b = make(filepath=path)
    assert not b.run(n_like_max=1500, n_eff=4000, discard_exploration=True)
    assert not b.explored and path.exists()
    c = make(filepath=path)
    assert c.n_like == b.n_like and len(c.bounds) == len(b.bounds)
    assert np.array_equal(c.shell_n, b.shell_n)
    assert np.allclose(c.shell_log_v, b.shell_log_v, rtol=0, atol=1e-12,
                       equal_nan=True)
    for pc, pb in zip(c.points, b.points):
        assert np.array_equal(pc, pb)
    assert np.array_equal(c.points_t, b.points_t)
    assert np.array_equal(c.shell_t, b.shell_t)
    # ... and once more after the exploration phase
    # (same trajectory as `a`, which was not done half-way)
    assert not c.run(n_like_max=a.n_like // 2, n_eff=4000,
                     discard_exploration=True)
    d = make(filepath=path)
    assert d.explored == c.explored and d.discard_exploration == \
        c.discard_exploration
    assert d.run(n_eff=4000, discard_exploration=True)

    assert d.n_like == a.n_like and len(d.bounds) == len(a.bounds)
    assert np.array_equal(d.shell_n, a.shell_n)
    assert abs(d.log_z - a.log_z) < 1e-9 and abs(d.n_eff - a.n_eff) < 1e-5
    pa, wa, la = a.posterior()
    pd, wd, ld = d.posterior()
    assert np.array_equal(pa, pd) and np.array_equal(la, ld)
    assert np.allclose(wa, wd, rtol=0, atol=1e-9)

    # the file of a finished run: nothing left to do
    e = make(filepath=path)
    assert e.run(n_eff=4000, discard_exploration=True)
    assert e.n_like == d.n_like and abs(e.log_z - d.log_z) < 1e-12
    # resume=False starts over
    f = make(filepath=path, resume=False)
    assert f.n_like == 0 and len(f.bounds) == 0


@pytest.mark.gpu
def test_resume_host_likelihood_with_blobs(tmp_path):
    """Python likelihood with blobs (sampler.py:886-906): exactly n_batch
    evaluations per step, the FIFOs of the bounds and the blobs are part of
    the state."""
    prior = Prior()
    prior.add_parameter('a', (-1, 1))
    prior.add_parameter('b', (-1, 1))

    def likelihood(p):
        log_l = -0.5 * ((p['a'] / 0.1)**2 + (p['b'] / 0.2)**2)
        return log_l, 2.0 * p['a'], (10 * p['b']).astype(np.int64)

    path = tmp_path / 'blobs.npz'

    def make(**kwargs):
        return Sampler(prior, likelihood, n_live=300, vectorized=True,
                       seed=5, **kwargs)

    a = make()
    assert a.run(n_eff=1500)
    b = make(filepath=path, checkpoint_interval=0.0)
    assert not b.run(n_like_max=900, n_eff=1500)
    c = make(filepath=path)
    assert c.run(n_eff=1500)
    assert c.n_like == a.n_like
    assert abs(c.log_z - a.log_z) < 1e-9
    pa, wa, la, ba = a.posterior(return_blobs=True)
    pc, wc, lc, bc = c.posterior(return_blobs=True)
    assert np.array_equal(pa, pc) and np.array_equal(ba, bc)
    assert np.allclose(bc['blob_0'], 2.0 * pc[:, 0])
    # write() refuses to clobber, write_shell_update() edits in place
    with pytest.raises(RuntimeError):
        c.write(path)
    with pytest.raises(ValueError):
        c.write(tmp_path / 'x.txt')
    c.write_shell_update(path, -1)
    with _store.open_store(path) as f:
        assert f['sampler'].attrs['n_like'] == c.n_like
        n_last = len(np.array(f['sampler']['log_l_{}'.format(
            len(c.bounds) - 1)]))
    assert n_last == len(c.log_l[-1])


def test_sampler_group_keys_are_the_references():
    """The attribute names Sampler.write stores in the 'sampler' group are
    the ones the reference's write() stores (parsed from its source), plus
    the documented extras; the dataset names follow the same patterns."""
    import ast
    import inspect
    from oracle import ref_arm
    if not ref_arm.available():
        pytest.skip('oracle/_ref missing (run oracle/make_ref.sh)')
    ref_arm._import_reference()
    import nautilus
    tree = ast.parse(inspect.getsource(nautilus.Sampler))
    keys = {}
    for fn in (n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)
               and n.name in ('write', 'write_shell_update')):
        found = []
        for node in ast.walk(fn):
            # `for key in [...]: group.attrs[key] = getattr(self, key)`
            if isinstance(node, ast.For) and isinstance(node.iter, ast.List):
                found += [e.value for e in node.iter.elts
                          if isinstance(e, ast.Constant)]
        keys[fn.name] = found
    mine = Sampler._ATTRS_CONFIG + Sampler._ATTRS_STATE
    assert set(keys['write']) - {'points_t', 'shell_t', 'log_l_t',
                                 'blobs_t'} == set(mine)
    src = inspect.getsource(Sampler.write_shell_update)
    for key in keys['write_shell_update']:
        assert "'{}'".format(key) in src, key
    src = inspect.getsource(Sampler.write) + inspect.getsource(
        Sampler._write_rng)
    for name in ('points_{}', 'log_l_{}', 'blobs_{}', 'points_t', 'shell_t',
                 'log_l_t', 'blobs_t', 'bound_{}', 'rng_state', 'rng_inc',
                 'rng_has_uint32', 'rng_uinteger', 'neural_network_{}'):
        assert name in src, name
