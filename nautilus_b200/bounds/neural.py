"""Neural-network bound: ellipsoid AND emulator score above a threshold.

Host-side mirror of ``nautilus/bounds/neural.py``.
"""

import numpy as np
import torch
from scipy.stats import rankdata

from .. import ops
from ..neural import NeuralNetworkEmulator
from .._device import default_device, to_device
from .basic import Ellipsoid, _DeviceBound


class NeuralBound(_DeviceBound):
    """(nautilus/bounds/neural.py:10-173)."""

    @classmethod
    def compute(cls, points, log_l, log_l_min, enlarge_per_dim=1.1,
                n_networks=4, neural_network_kwargs={}, pool=None, rng=None,
                mode=None, defer=False):
        """``defer=True`` returns as soon as the fit kernel is enqueued;
        ``finish()`` must be called before the bound is used (the caller may
        build other bounds in between, on another stream)."""
        bound = cls()
        bound.mode = NeuralNetworkEmulator.mode if mode is None else mode
        on_device = isinstance(points, torch.Tensor)
        if not on_device:
            points = np.asarray(points, dtype=float)
            log_l = np.asarray(log_l, dtype=float)
        bound.n_dim = points.shape[1]
        if rng is None:
            rng = np.random.default_rng()

        live = log_l >= log_l_min
        live_points = points[live]
        if on_device:
            live_points = live_points.cpu().numpy()
        bound.outer_bound = Ellipsoid.compute(
            live_points, enlarge_per_dim=enlarge_per_dim, rng=rng)
        if n_networks == 0:
            bound.emulator = None
            bound.score_predict_min = 0
            return bound

        # training set: every known point inside the enlarged live ellipsoid,
        # in its whitened frame; target = rank score, live points in (1/2, 1),
        # the others in (0, 1/2)   (neural.py:78-88)
        inside = bound.outer_bound.contains(points)
        points, log_l = points[inside], log_l[inside]
        whitened = bound.outer_bound.transform(points)
        if on_device:        # the training subset (~2 n_live rows) only
            whitened = whitened.cpu().numpy()
            log_l = log_l.cpu().numpy()
        live = log_l >= log_l_min
        score = np.empty(len(points))
        score[live] = 0.5 + 0.5 * (rankdata(log_l[live]) - 0.5) / np.sum(live)
        score[~live] = 0.5 * (rankdata(log_l[~live]) - 0.5) / max(
            np.sum(~live), 1)
        bound.emulator = NeuralNetworkEmulator.train_async(
            whitened, score, n_networks=n_networks,
            neural_network_kwargs=neural_network_kwargs, pool=pool,
            seed=int(rng.integers(0, 2**63 - 1)))
        bound._unfinished = (whitened, score, live)
        return bound if defer else bound.finish()

    def finish(self):
        """Second half of ``compute``: wait for the fit, set the threshold."""
        unfinished = getattr(self, '_unfinished', None)
        if unfinished is None:
            return self
        whitened, score, live = unfinished
        self._unfinished = None
        self.emulator.wait()
        # threshold: cubic fit of predicted vs true score, evaluated at the
        # lowest live score (neural.py:93-95)
        self.emulator.mode = self.mode
        predicted = self.emulator.predict(whitened)
        self.score_predict_min = float(np.polyval(
            np.polyfit(score, predicted, 3), np.amin(score[live])))
        return self

    def write(self, group):
        """(nautilus/bounds/neural.py:128-141)."""
        group.attrs['n_dim'] = self.n_dim
        group.attrs['score_predict_min'] = self.score_predict_min
        self.outer_bound.write(group.create_group('outer_bound'))
        if self.emulator is not None:
            self.emulator.write(group.create_group('emulator'))

    @classmethod
    def read(cls, group, rng=None, mode=None):
        """(nautilus/bounds/neural.py:143-173).  ``mode``: emulator
        arithmetic of ``contains`` (as in ``compute``)."""
        bound = cls()
        bound.mode = NeuralNetworkEmulator.mode if mode is None else mode
        bound.n_dim = int(group.attrs['n_dim'])
        bound.score_predict_min = float(group.attrs['score_predict_min'])
        bound.outer_bound = Ellipsoid.read(group['outer_bound'], rng=rng)
        bound.emulator = (NeuralNetworkEmulator.read(group['emulator'])
                          if 'emulator' in group else None)
        return bound

    def nb_spec(self):
        return dict(ell=self.outer_bound.ell_spec(),
                    emulator=None if self.emulator is None
                    else self.emulator.emu_spec(),
                    score_predict_min=float(self.score_predict_min))

    def spec(self):
        d = self.n_dim
        everything = dict(c=np.zeros(d), B=np.eye(d) * 1e300,
                          B_inv=np.eye(d) * 1e-300)
        return dict(kind='nautilus', n_dim=d, unit=False,
                    log_v_all=np.zeros(1),
                    mixtures=[dict(dim_cube=np.zeros(d, bool),
                                   ell=everything)],
                    neural=[self.nb_spec()])

    def contains(self, points, mode=None):
        """Inside the ellipsoid and score above the threshold
        (neural.py:99-126)."""
        t, restore = to_device(points, self.n_dim)
        mode = getattr(self, 'mode', NeuralNetworkEmulator.mode) \
            if mode is None else mode
        return restore(self._device_stack().contains(0, t, which=2,
                                                     mode=mode))
