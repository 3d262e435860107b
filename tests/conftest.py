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


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz')) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture
def golden():
    return load_golden
