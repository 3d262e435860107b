"""Basic bounds: unit cube, ellipsoid, cube-ellipsoid mixture.

Host-side mirror of ``nautilus/bounds/basic.py``: the same classes, methods,
argument meaning and errors; parameters live in NumPy on the host exactly as
the reference stores them (``c``, ``A``, ``B``, ``B_inv``) and every
``contains`` / ``transform`` / ``sample`` runs on the GPU through the C ABI
(nautilus_b200/ops.py).  ``points`` may be NumPy arrays (results come back as
NumPy, drop-in for the reference) or CUDA tensors (results stay on device).
"""

import numpy as np
import torch
from scipy.special import gammaln

from .. import ops
from . import _construct
from .._device import PhiloxStream, default_device, to_device


class _DeviceBound:
    """Lazily serialises ``self.spec()`` to the device."""

    _stack = None

    def _device_stack(self):
        if self._stack is None:
            self._stack = ops.DeviceStack([self.spec()],
                                          device=default_device())
        return self._stack

    def _invalidate(self):
        self._stack = None

    def _propose(self, n):
        """n raw draws from this bound's own stream: (points, code)."""
        stack = self._device_stack()
        offset = self.stream.take(n)
        points, code, _ = stack.propose(0, int(n), seed=self.stream.seed,
                                        offset=offset,
                                        stream_id=self.stream.stream_id)
        return points, code

    def reset(self, rng=None):
        """Reset random number generation (basic.py:139-151)."""
        if rng is not None:
            self.stream.reseed(rng)


class UnitCube(_DeviceBound):
    """Unit hypercube 0 <= x_i < 1 (nautilus/bounds/basic.py:9-151)."""

    @classmethod
    def compute(cls, n_dim, rng=None):
        bound = cls()
        bound.n_dim = int(n_dim)
        bound.stream = PhiloxStream(rng)
        return bound

    def spec(self):
        return dict(kind='cube', n_dim=self.n_dim)

    def contains(self, points):
        t, restore = to_device(points, self.n_dim)
        return restore(self._device_stack().contains(0, t))

    def sample(self, n_points=100, pool=None, as_numpy=True):
        points, _ = self._propose(n_points)
        return points.cpu().numpy() if as_numpy else points

    @property
    def log_v(self):
        return 0

    def write(self, group):
        """(basic.py:100-110)."""
        group.attrs['type'] = 'UnitCube'
        group.attrs['n_dim'] = self.n_dim
        self.stream.write(group)

    @classmethod
    def read(cls, group, rng=None):
        """(basic.py:112-137)."""
        bound = cls()
        bound.n_dim = int(group.attrs['n_dim'])
        bound.stream = PhiloxStream.read(group, rng)
        return bound


class Ellipsoid(_DeviceBound):
    r"""Ellipsoid (x - c)^T A (x - c) <= 1 (nautilus/bounds/basic.py:244-449).

    ``B`` is the Cholesky factor of ``A^{-1}`` and ``B_inv`` its inverse.
    """

    @classmethod
    def compute(cls, points, enlarge_per_dim=1.1, rng=None):
        points = np.asarray(points, dtype=float)
        if enlarge_per_dim < 1.0:
            raise ValueError(
                "The 'enlarge_per_dim' factor cannot be smaller than unity.")
        if not points.shape[0] > points.shape[1]:
            raise ValueError('Number of points must be larger than number '
                             'dimensions.')
        c, a, a_inv = _construct.enclosing_ellipsoid(
            points, device=_construct.construction_device())
        return cls.from_matrices(c, a / enlarge_per_dim**2.0,
                                 a_inv * enlarge_per_dim**2.0, rng=rng)

    @classmethod
    def from_matrices(cls, c, A, A_inv, rng=None):
        bound = cls()
        bound.n_dim = len(c)
        bound.c = np.array(c, dtype=float)
        bound.A = np.array(A, dtype=float)
        bound.B = np.linalg.cholesky(A_inv)
        # triangular solve: B_inv is exactly lower-triangular, so the packed
        # fast path of the kernels applies
        bound.B_inv = np.tril(np.linalg.solve(bound.B, np.eye(bound.n_dim)))
        bound.stream = PhiloxStream(rng)
        return bound

    def ell_spec(self):
        return dict(c=self.c, B=self.B, B_inv=self.B_inv)

    def spec(self):
        return dict(kind='nautilus', n_dim=self.n_dim, unit=False,
                    log_v_all=np.array([self.log_v]),
                    mixtures=[dict(dim_cube=np.zeros(self.n_dim, dtype=bool),
                                   ell=self.ell_spec())], neural=[])

    def _params(self):
        if getattr(self, '_dev_params', None) is None:
            dev = default_device()
            self._dev_params = tuple(torch.from_numpy(
                np.ascontiguousarray(a)).to(dev) for a in
                (self.c, self.B, self.B_inv))
        return self._dev_params

    def transform(self, points, inverse=False):
        t, restore = to_device(points, self.n_dim)
        c, B, B_inv = self._params()
        return restore(ops.ell_transform(t, c, B if inverse else B_inv,
                                         inverse=inverse))

    def contains(self, points):
        t, restore = to_device(points, self.n_dim)
        c, _, B_inv = self._params()
        return restore(ops.ell_contains(t, c, B_inv))

    def sample(self, n_points=100, as_numpy=True):
        points, _ = self._propose(n_points)
        return points.cpu().numpy() if as_numpy else points

    @property
    def log_v(self):
        """log volume = log|det B| + (d/2) log(pi) - lgamma(d/2 + 1)
        (basic.py:393-394)."""
        return (np.sum(np.log(np.diag(self.B))) +
                0.5 * self.n_dim * np.log(np.pi) -
                gammaln(self.n_dim / 2.0 + 1))

    def write(self, group):
        """(basic.py:396-407)."""
        group.attrs['type'] = 'Ellipsoid'
        for key in ['n_dim', 'c', 'A', 'B', 'B_inv']:
            group.attrs[key] = getattr(self, key)
        self.stream.write(group)

    @classmethod
    def read(cls, group, rng=None):
        """(basic.py:409-437).  The matrices are taken as stored (a file
        written by the reference holds a dense ``B_inv``; the kernels accept
        both forms)."""
        bound = cls()
        bound.n_dim = int(group.attrs['n_dim'])
        for key in ['c', 'A', 'B', 'B_inv']:
            setattr(bound, key, np.array(group.attrs[key], dtype=float))
        bound.stream = PhiloxStream.read(group, rng)
        return bound


class UnitCubeEllipsoidMixture(_DeviceBound):
    """Some dimensions bounded by the unit range, the others by an ellipsoid
    (nautilus/bounds/basic.py:452-726)."""

    @classmethod
    def compute(cls, points, enlarge_per_dim=1.1, rng=None):
        points = np.asarray(points, dtype=float)
        n_dim = points.shape[1]
        kwargs = dict(enlarge_per_dim=enlarge_per_dim, rng=rng)

        # dimension search (basic.py:497-512): one [n, m] x [m, m] product on
        # the construction device gives the projected volume of EVERY
        # candidate (_construct.projection_scan); the host only reads the
        # argmin
        dev = _construct.construction_device()
        pts_t = torch.from_numpy(np.ascontiguousarray(points))
        if dev is not None:
            pts_t = pts_t.to(dev)

        def best_drop(ell, cols):
            trial = _construct.projection_scan(
                pts_t[:, cols],
                torch.from_numpy(ell.c).to(pts_t.device),
                torch.from_numpy(ell.A).to(pts_t.device))
            return int(torch.argmin(trial).item())

        # backward pass: hand dimensions to the cube while the volume shrinks
        in_cube = np.zeros(n_dim, dtype=bool)
        ell = Ellipsoid.compute(points, **kwargs)
        while np.sum(~in_cube) > 1:
            cols = list(np.flatnonzero(~in_cube))
            cand = cols[best_drop(ell, cols)]
            in_cube[cand] = True
            smaller = Ellipsoid.compute(points[:, ~in_cube], **kwargs)
            if smaller.log_v < ell.log_v:
                ell = smaller
            else:
                in_cube[cand] = False
                break

        # forward pass if even that ellipsoid is larger than the cube
        if ell.log_v > 0:
            ell = None
            in_cube = np.ones(n_dim, dtype=bool)
            best = 0.0
            improved = True
            while improved:
                improved = False
                for dim in np.flatnonzero(in_cube):
                    trial_mask = in_cube.copy()
                    trial_mask[dim] = False
                    if np.sum(~trial_mask) < 1:
                        continue
                    cand = Ellipsoid.compute(points[:, ~trial_mask], **kwargs)
                    if cand.log_v < best:
                        best, ell, in_cube = cand.log_v, cand, trial_mask
                        improved = True

        bound = cls()
        bound.n_dim = n_dim
        bound.dim_cube = in_cube
        bound.cube = (UnitCube.compute(int(np.sum(in_cube)), rng=rng)
                      if np.any(in_cube) else None)
        bound.ellipsoid = None if np.all(in_cube) else ell
        bound.stream = PhiloxStream(rng)
        return bound

    def mix_spec(self):
        return dict(dim_cube=self.dim_cube,
                    ell=None if self.ellipsoid is None
                    else self.ellipsoid.ell_spec())

    def spec(self):
        return dict(kind='nautilus', n_dim=self.n_dim, unit=False,
                    log_v_all=np.array([float(self.log_v)]),
                    mixtures=[self.mix_spec()], neural=[])

    def transform(self, points):
        """Cube dims -> [-1, 1], ellipsoid dims -> whitened (basic.py:565-592).
        """
        t, restore = to_device(points, self.n_dim)
        out = t.clone()
        cube = torch.from_numpy(self.dim_cube).to(t.device)
        if self.cube is not None:
            out[:, cube] = t[:, cube] * 2 - 1
        if self.ellipsoid is not None:
            out[:, ~cube] = self.ellipsoid.transform(t[:, ~cube].contiguous())
        return restore(out)

    def contains(self, points):
        t, restore = to_device(points, self.n_dim)
        count, _ = self._device_stack().union_count(0, t)
        return restore(count > 0)

    def sample(self, n_points=100, as_numpy=True):
        points, _ = self._propose(n_points)
        return points.cpu().numpy() if as_numpy else points

    @property
    def log_v(self):
        return 0 if self.ellipsoid is None else self.ellipsoid.log_v

    def write(self, group):
        """(basic.py:657-673)."""
        group.attrs['type'] = 'UnitCubeEllipsoidMixture'
        group.attrs['n_dim'] = self.n_dim
        group.create_dataset('dim_cube', data=self.dim_cube)
        if self.cube is not None:
            self.cube.write(group.create_group('cube'))
        if self.ellipsoid is not None:
            self.ellipsoid.write(group.create_group('ellipsoid'))
        self.stream.write(group)

    @classmethod
    def read(cls, group, rng=None):
        """(basic.py:675-712)."""
        bound = cls()
        bound.n_dim = int(group.attrs['n_dim'])
        bound.dim_cube = np.array(group['dim_cube'], dtype=bool)
        bound.cube = (UnitCube.read(group['cube'], rng=rng)
                      if np.any(bound.dim_cube) else None)
        bound.ellipsoid = (Ellipsoid.read(group['ellipsoid'], rng=rng)
                           if not np.all(bound.dim_cube) else None)
        bound.stream = PhiloxStream.read(group, rng)
        return bound

    def reset(self, rng=None):
        if rng is not None:
            self.stream.reseed(rng)
            if self.ellipsoid is not None:
                self.ellipsoid.reset(rng)
            if self.cube is not None:
                self.cube.reset(rng)
