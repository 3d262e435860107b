"""Union of ellipsoids / cube-ellipsoid mixtures.

Host-side mirror of ``nautilus/bounds/union.py``.  Sampling keeps the
reference's semantics -- volume-proportional choice, unit-cube cut, accept
with probability 1/n_overlap, integer ``n_sample`` / ``n_reject`` counters
that feed the volume estimate, a FIFO of accepted points -- but draws whole
batches on the GPU (csrc/nb200_kernels.cu:k_union_propose).
"""

import numpy as np
import torch
from scipy.special import logsumexp

from .. import ops
from . import _construct
from .._device import PhiloxStream, default_device, to_device
from .basic import Ellipsoid, UnitCube, UnitCubeEllipsoidMixture, _DeviceBound

ellipsoids_overlap = _construct.ellipsoids_overlap


class Union(_DeviceBound):
    """Union of K bounds, optionally cut by the unit cube
    (nautilus/bounds/union.py:43-450)."""

    raw_batch = 1 << 16      # raw draws per kernel launch (reference: 1000)

    @classmethod
    def compute(cls, points, enlarge_per_dim=1.1, n_points_min=None,
                unit=True, bound_class=Ellipsoid, rng=None):
        points = np.asarray(points, dtype=float)
        bound = cls()
        bound.n_dim = points.shape[1]
        bound.enlarge_per_dim = enlarge_per_dim
        if n_points_min is None:
            n_points_min = bound.n_dim + 1
        elif n_points_min < bound.n_dim + 1:
            raise ValueError('The number of points per bound must be '
                             'larger than the number of dimensions.')
        bound.n_points_min = n_points_min
        bound.cube = UnitCube.compute(bound.n_dim, rng=rng) if unit else None
        bound.bound_class = bound_class
        bound.rng = np.random.default_rng() if rng is None else rng
        bound.points_bounds = [points]
        bound.bounds = [bound_class.compute(
            points, enlarge_per_dim=enlarge_per_dim, rng=rng)]
        bound.log_v_all = np.array([bound.bounds[0].log_v], dtype=float)
        bound.block = np.atleast_1d(len(points) < 2 * n_points_min)
        bound.stream = PhiloxStream(rng)
        bound._clear()
        return bound

    # -- state -------------------------------------------------------------
    def _clear(self):
        self.points = np.zeros((0, self.n_dim))
        self._buffer = None          # device FIFO of accepted points
        self.n_sample = 0
        self.n_reject = 0
        self._invalidate()

    def spec(self):
        mixtures = []
        for b in self.bounds:
            if isinstance(b, Ellipsoid):
                mixtures.append(dict(dim_cube=np.zeros(self.n_dim, bool),
                                     ell=b.ell_spec()))
            else:
                mixtures.append(b.mix_spec())
        return dict(kind='nautilus', n_dim=self.n_dim,
                    unit=self.cube is not None,
                    log_v_all=np.asarray(self.log_v_all, dtype=float),
                    mixtures=mixtures, neural=[])

    # -- construction ---------------------------------------------------------
    def split(self, allow_overlap=True):
        """Split the largest splittable bound in two (union.py:153-229)."""
        if not allow_overlap and not isinstance(self.bounds[0], Ellipsoid):
            raise ValueError("'allow_overlap' can only be False if "
                             "bounds are ellipsoids.")
        while True:
            if np.all(self.block):
                return False
            index = int(np.argmax(np.where(~self.block, self.log_v_all,
                                           -np.inf)))
            pts = self.points_bounds[index]
            whitened = self.bounds[index].transform(pts)
            # EM of all restarts advanced together on the device
            dev = _construct.construction_device()
            gmm_rng = np.random.default_rng(self.rng.integers(2**32 - 1))
            if dev is None:
                log_p = _construct.two_gaussians(whitened, gmm_rng)
            else:
                log_p = _construct.two_gaussians_batched(whitened, gmm_rng,
                                                         device=dev)
            labels = np.argmax(log_p, axis=1)
            counts = np.bincount(labels, minlength=2)
            if np.any(counts < self.n_points_min):
                # top up the smaller cluster with its most likely members
                small = int(np.argmin(counts))
                labels[np.argsort(-log_p[:, small])[:self.n_points_min]] = \
                    small
            halves = [pts[labels == k] for k in (0, 1)]
            try:
                new = [type(self.bounds[0]).compute(
                    h, enlarge_per_dim=self.enlarge_per_dim, rng=self.rng)
                    for h in halves]
            except (ValueError, np.linalg.LinAlgError):
                self.block[index] = True
                continue
            others = self.bounds[:index] + self.bounds[index + 1:]
            if not allow_overlap and ellipsoids_overlap(
                    others + new, device=_construct.construction_device()):
                return False
            if logsumexp([new[0].log_v, new[1].log_v]) > \
                    self.bounds[index].log_v:
                self.block[index] = True
                continue
            self.points_bounds = (self.points_bounds[:index] +
                                  self.points_bounds[index + 1:] + halves)
            self.bounds = others + new
            self.log_v_all = np.array([b.log_v for b in self.bounds], float)
            self.block = np.concatenate(
                [np.delete(self.block, index),
                 [len(h) < 2 * self.n_points_min for h in halves]])
            self._clear()
            return True

    def trim(self, threshold=1e3):
        """Drop the lowest-density bound if it is `threshold` times less
        dense than the median of the others (union.py:231-267)."""
        if len(self.bounds) == 1:
            return False
        log_density = np.array(
            [np.log(len(p)) - b.log_v
             for p, b in zip(self.points_bounds, self.bounds)])
        index = int(np.argmin(log_density))
        rest = np.delete(log_density, index)
        if log_density[index] - np.median(rest) >= -np.log(threshold):
            return False
        del self.points_bounds[index], self.bounds[index]
        self.block = np.delete(self.block, index)
        self.log_v_all = np.array([b.log_v for b in self.bounds], float)
        self._clear()
        return True

    # -- device operations ---------------------------------------------------
    def contains(self, points):
        t, restore = to_device(points, self.n_dim)
        _, inside = self._device_stack().union_count(0, t)
        return restore(inside)

    def _refill(self, n_points):
        have = 0 if self._buffer is None else self._buffer.shape[0]
        chunks = [] if self._buffer is None else [self._buffer]
        while have < n_points:
            acc = 1.0 - self.n_reject / self.n_sample if self.n_sample else 0.5
            n_raw = int(min(max(self.raw_batch,
                                1.2 * (n_points - have) / max(acc, 1e-3)),
                            1 << 22))
            pts, code = self._propose(n_raw)
            keep = pts[code == ops.CODE_IN_SHELL]
            chunks.append(keep)
            have += keep.shape[0]
            self.n_sample += n_raw
            self.n_reject += n_raw - keep.shape[0]
        self._buffer = torch.cat(chunks) if len(chunks) > 1 else chunks[0]

    def sample(self, n_points=100, as_numpy=True):
        """Pop n accepted points from the FIFO, refilling it with raw batches
        (union.py:291-327)."""
        self._refill(n_points)
        out = self._buffer[:n_points]
        self._buffer = self._buffer[n_points:]
        return out.cpu().numpy() if as_numpy else out.contiguous()

    @property
    def log_v(self):
        """log sum_k V_k + log(1 - n_reject / n_sample) (union.py:329-343)."""
        if self.n_sample == 0:
            self._refill(100)
        return logsumexp(self.log_v_all) + np.log(
            1.0 - self.n_reject / self.n_sample)

    # -- checkpoints --------------------------------------------------------
    def _fifo_host(self):
        """The FIFO of accepted points the reference stores as ``points``."""
        if self._buffer is None:
            return np.zeros((0, self.n_dim))
        return self._buffer.cpu().numpy()

    def write(self, group):
        """Layout of nautilus/bounds/union.py:345-370 (type 'MultiEllipsoid'
        is the reference's name for it), plus the split bookkeeping and the
        Philox stream."""
        group.attrs['type'] = 'MultiEllipsoid'
        for key in ['n_dim', 'log_v_all', 'enlarge_per_dim', 'n_points_min',
                    'n_sample', 'n_reject']:
            group.attrs[key] = getattr(self, key)
        group.attrs['unit'] = self.cube is not None
        if self.cube is not None:
            self.cube.write(group.create_group('cube'))
        group.attrs['bound_class'] = self.bounds[0].__class__.__name__
        group.attrs['block'] = np.asarray(self.block, dtype=bool)
        for i, bound in enumerate(self.bounds):
            bound.write(group.create_group('bound_{}'.format(i)))
        for i, points in enumerate(self.points_bounds):
            group.create_dataset('points_bound_{}'.format(i), data=points)
        group.create_dataset('points', data=self._fifo_host(),
                             maxshape=(None, self.n_dim))
        self.stream.write(group)

    def update(self, group):
        """(union.py:372-384)."""
        group.attrs['n_sample'] = self.n_sample
        group.attrs['n_reject'] = self.n_reject
        fifo = self._fifo_host()
        group['points'].resize(fifo.shape)
        group['points'][...] = fifo
        self.stream.write(group)

    @classmethod
    def read(cls, group, rng=None):
        """(union.py:386-429)."""
        bound = cls()
        bound.rng = np.random.default_rng() if rng is None else rng
        bound.n_dim = int(group.attrs['n_dim'])
        bound.log_v_all = np.atleast_1d(np.array(group.attrs['log_v_all'],
                                                 dtype=float))
        bound.enlarge_per_dim = float(group.attrs['enlarge_per_dim'])
        bound.n_points_min = int(group.attrs['n_points_min'])
        bound.cube = (UnitCube.read(group['cube'], rng=bound.rng)
                      if group.attrs['unit'] else None)
        bound.bound_class = (Ellipsoid if group.attrs['bound_class'] ==
                             'Ellipsoid' else UnitCubeEllipsoidMixture)
        n = len(bound.log_v_all)
        bound.bounds = [bound.bound_class.read(
            group['bound_{}'.format(i)], rng=bound.rng) for i in range(n)]
        bound.points_bounds = [np.array(group['points_bound_{}'.format(i)])
                               for i in range(n)]
        bound.block = (np.array(group.attrs['block'], dtype=bool)
                       if 'block' in group.attrs else np.array(
                           [len(p) < 2 * bound.n_points_min
                            for p in bound.points_bounds]))
        bound.stream = PhiloxStream.read(group, bound.rng)
        bound._clear()
        bound.n_sample = int(group.attrs['n_sample'])
        bound.n_reject = int(group.attrs['n_reject'])
        fifo = np.array(group['points'], dtype=float)
        if len(fifo) > 0:
            bound._buffer = torch.from_numpy(
                np.ascontiguousarray(fifo)).to(default_device())
        return bound

    def reset(self, rng=None):
        """Forget sampling progress; optionally reseed (union.py:431-450)."""
        self._buffer = None
        self.points = np.zeros((0, self.n_dim))
        self.n_sample = 0
        self.n_reject = 0
        if rng is not None:
            self.rng = rng
            self.stream.reseed(rng)
            for b in self.bounds:
                b.reset(rng)
