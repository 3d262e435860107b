"""pytest configuration: markers, paths and shared fixtures."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: test needs a CUDA device (run on the B200 box)')


def _gpu_skip_reason():
    """None if the `gpu` tests can run here, else why not."""
    try:
        import torch
        if not torch.cuda.is_available():
            return 'needs a CUDA device'
    except ImportError:
        return 'needs torch with CUDA'
    from nautilus_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        return 'libnautilus_b200.so is not built'
    return None


def pytest_collection_modifyitems(config, items):
    """Skip (not fail) the `gpu` tests on a host that cannot run them, so a
    plain `pytest tests` on a CPU box shows host-side regressions only."""
    reason = None
    for item in items:
        if 'gpu' in item.keywords:
            if reason is None:
                reason = _gpu_skip_reason() or ''
            if reason:
                item.add_marker(pytest.mark.skip(reason=reason))


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz')) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture
def golden():
    return load_golden
