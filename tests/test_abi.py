"""The C-ABI library builds, loads and exports every symbol the public header
declares (no compute calls: this runs without a GPU)."""

import os
import re
import subprocess

import pytest

from nautilus_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'nautilus_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(nb200_\w+)\s*\(', text)))


@pytest.fixture(scope='module')
def built():
    _lib.build()
    return _lib.LIB_PATH


def test_header_and_prototypes_agree():
    assert header_symbols() == sorted(_lib.PROTOTYPES)


def test_library_exports_every_symbol(built):
    out = subprocess.check_output(['nm', '-D', '--defined-only', built],
                                  text=True)
    exported = set(re.findall(r' T (nb200_\w+)', out))
    assert set(header_symbols()) <= exported


def test_library_loads_and_reports_version(built):
    lib = _lib.lib()
    assert lib.nb200_version() == 200
    assert lib.nb200_workspace_bytes(1 << 20, 30) > (1 << 20) * 30 * 8


def test_library_is_sm100a_only(built):
    out = subprocess.check_output(
        ['/usr/local/cuda/bin/cuobjdump', '-lelf', built], text=True)
    archs = set(re.findall(r'sm_(\w+)\.', out))
    assert archs == {'100a'}, archs


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_lib.NautilusB200Error):
        _lib.lib()


def test_session_fails_loudly_without_gpu(built):
    # no CPU fallback anywhere: on a box without a CUDA device the host-buffer
    # session refuses to exist instead of computing something on the host
    import numpy as np
    torch = pytest.importorskip('torch')
    if torch.cuda.is_available():
        pytest.skip('needs a box without a GPU')
    from nautilus_b200 import ops
    spec = dict(kind='cube', n_dim=3)
    with pytest.raises(_lib.NautilusB200Error, match='CUDA error'):
        ops.HostSession([spec], n_max=1024)
    # argument checks come before any CUDA call
    with pytest.raises(_lib.NautilusB200Error, match='n_max'):
        ops.HostSession([spec], n_max=0)
    assert np.isfinite(_lib.lib().nb200_workspace_bytes(1024, 3))


def test_torch_ops_registered_through_the_stable_abi(built):
    """torch.ops.nautilus_b200.* (STABLE_TORCH_LIBRARY shim over the C ABI):
    builds, loads, registers its schemas; CUDA-only, so a CPU call raises
    instead of computing anything on the host."""
    torch = pytest.importorskip('torch')
    _lib.build_torch_ops()
    ops = _lib.torch_ops()
    for name in ('ell_contains', 'shell_stats', 'bound_contains',
                 'shell_cycle'):
        assert hasattr(ops, name)
    schema = str(torch.ops.nautilus_b200.shell_cycle.default._schema)
    assert 'like_params' in schema and '-> (Tensor, Tensor, Tensor' in schema
    out = subprocess.check_output(['nm', '-D', '--undefined-only',
                                   _lib.TORCH_LIB_PATH], text=True)
    assert 'dlsym' in out and 'aoti_torch_get_current_cuda_stream' in out
    # stable ABI only: no ATen / c10 C++ symbols are pulled in
    assert not re.search(r' U _ZN2at|_ZN3c10[^d]', out)
    with pytest.raises(NotImplementedError):
        ops.ell_contains(torch.zeros((3, 2), dtype=torch.float64),
                         torch.zeros(2, dtype=torch.float64),
                         torch.eye(2, dtype=torch.float64))
