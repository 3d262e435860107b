"""ctypes loader of the canonical-order C oracle (oracle/c/nb200_oracle.c).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'c')
_LIB = None


def build():
    subprocess.check_call(['make', '-C', _HERE, '-s'])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libnb200_oracle.so')
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ell_transform(points, c, M, inverse=False):
    points = _f64(points)
    n, d = points.shape
    out = np.empty_like(points)
    lib().orc_ell_transform(_p(points), ctypes.c_int64(n), d, _p(_f64(c)),
                            _p(_f64(M)), int(inverse), _p(out))
    return out


def ell_contains(points, c, B_inv):
    points = _f64(points)
    n, d = points.shape
    out = np.empty(n, dtype=np.uint8)
    r2 = np.empty(n)
    lib().orc_ell_contains(_p(points), ctypes.c_int64(n), d, _p(_f64(c)),
                           _p(_f64(B_inv)), _p(out), _p(r2))
    return out.astype(bool), r2


def union_count(meta, data, bound, points):
    points = _f64(points)
    n = len(points)
    count = np.empty(n, dtype=np.int32)
    contains = np.empty(n, dtype=np.uint8)
    lib().orc_union_count(_p(meta), _p(data), bound, _p(points),
                          ctypes.c_int64(n), _p(count), _p(contains))
    return count, contains.astype(bool)


def neural(meta, data, bound, j, points):
    points = _f64(points)
    n, d = points.shape
    in_ell = np.empty(n, dtype=np.uint8)
    t_rows = np.empty((n, d))
    score = np.empty(n)
    ok = np.empty(n, dtype=np.uint8)
    lib().orc_neural(_p(meta), _p(data), bound, j, _p(points),
                     ctypes.c_int64(n), _p(in_ell), _p(t_rows), _p(score),
                     _p(ok))
    return in_ell.astype(bool), t_rows, score, ok.astype(bool)


def mlp_predict(meta, data, bound, j, t_rows):
    t_rows = _f64(t_rows)
    out = np.empty(len(t_rows))
    lib().orc_mlp_predict(_p(meta), _p(data), bound, j, _p(t_rows),
                          ctypes.c_int64(len(t_rows)), _p(out))
    return out


def bound_contains(meta, data, bound, points):
    points = _f64(points)
    out = np.empty(len(points), dtype=np.uint8)
    lib().orc_bound_contains(_p(meta), _p(data), bound, _p(points),
                             ctypes.c_int64(len(points)), _p(out))
    return out.astype(bool)
