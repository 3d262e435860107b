"""Importance nested sampling driver.

Drop-in for ``nautilus.Sampler`` (nautilus/sampler.py): same constructor and
``run`` / ``posterior`` signatures, same attributes (``bounds``, ``points``,
``log_l``, ``shell_*``), same control flow between shells -- with the
per-shell proposal -> neural filter -> exclusion -> likelihood ->
importance-weight cycle executed by the CUDA kernels behind
``include/nautilus_b200.h``.  The control plane (this file) is host Python
like the reference's; stored samples are NumPy arrays on the host so user
code that reads ``sampler.points`` keeps working.

Differences that are deliberate and documented in DESIGN.md:
  * random numbers come from Philox streams seeded by ``seed`` (the reference
    threads one PCG64 generator through everything), so runs are reproducible
    but not draw-for-draw identical to the reference;
  * ``likelihood`` may be a ``nautilus_b200.likelihoods.DeviceLikelihood``;
    it is then evaluated on the GPU on unit-cube points;
  * HDF5 checkpointing and periodic parameters are out of scope and raise.
"""

from functools import partial
from shutil import get_terminal_size
from time import time

import numpy as np
import torch
from scipy.special import logsumexp

from . import ops
from .bounds import NautilusBound, UnitCube
from ._device import default_device
from .likelihoods import DeviceLikelihood
from .pool import GpuPool, NautilusPool, likelihood_worker


class Sampler:
    """Importance nested sampler (nautilus/sampler.py:21-1377)."""

    def __init__(self, prior, likelihood, n_dim=None, n_live=2000,
                 n_update=None, enlarge_per_dim=1.1, n_points_min=None,
                 split_threshold=100, periodic=None, n_networks=4,
                 neural_network_kwargs={}, prior_args=[], prior_kwargs={},
                 likelihood_args=[], likelihood_kwargs={}, n_batch=None,
                 n_like_new_bound=None, vectorized=False, pass_dict=None,
                 pool=None, seed=None, blobs_dtype=None, filepath=None,
                 resume=True, emulator_arith='auto'):
        if filepath is not None:
            raise NotImplementedError(
                'HDF5 checkpointing is outside the scope of nautilus_b200.')
        if periodic is not None:
            raise NotImplementedError(
                'periodic parameters are outside the scope of nautilus_b200.')

        self.device_likelihood = isinstance(likelihood, DeviceLikelihood)
        if callable(prior):
            self.prior = partial(prior, *prior_args, **prior_kwargs)
            if n_dim is None:
                raise ValueError("When passing a function as the 'prior' "
                                 "argument, 'n_dim' cannot be None.")
            self.n_dim = n_dim
            pass_dict = False if pass_dict is None else pass_dict
        else:
            self.prior = prior
            self.n_dim = prior.dimensionality()
            pass_dict = True if pass_dict is None else pass_dict
        if self.device_likelihood:
            self.likelihood = likelihood
        else:
            self.likelihood = partial(
                likelihood, *likelihood_args, **likelihood_kwargs)
        if self.n_dim <= 1:
            raise ValueError(
                'Cannot run Nautilus with less than 2 parameters.')

        self.n_live = n_live
        self.n_update = n_live if n_update is None else n_update
        self.n_like_new_bound = (10 * n_live if n_like_new_bound is None
                                 else n_like_new_bound)
        self.enlarge_per_dim = enlarge_per_dim
        self.n_points_min = (self.n_dim + 50 if n_points_min is None
                             else n_points_min)
        self.split_threshold = split_threshold
        self.periodic = None
        self.n_networks = n_networks
        self.neural_network_kwargs = neural_network_kwargs
        self.vectorized = vectorized
        self.pass_dict = pass_dict
        # 'auto': tensor cores (tf32) whenever the emulator architecture fits
        # the tcgen05 kernel, fp64 otherwise; 'f64' is the bit-parity mode
        self.emulator_arith = emulator_arith
        self.mlp_mode = {'f64': ops.MLP_F64, 'tf32': ops.MLP_TF32,
                         'auto': ops.MLP_F64}[emulator_arith]
        if emulator_arith == 'auto':
            from ._pack import pack_tc
            hidden = tuple(np.atleast_1d(neural_network_kwargs.get(
                'hidden_layer_sizes', (100, 50, 20))))
            sizes = (self.n_dim, ) + tuple(int(h) for h in hidden) + (1, )
            probe = dict(
                coefs=[[np.zeros((a, b)) for a, b in zip(sizes[:-1],
                                                         sizes[1:])]],
                intercepts=[[np.zeros(b) for b in sizes[1:]]])
            probe['coefs'] = probe['coefs'] * max(n_networks, 1)
            probe['intercepts'] = probe['intercepts'] * max(n_networks, 1)
            if n_networks > 0 and pack_tc(probe, 0.0) is not None:
                self.mlp_mode = ops.MLP_TF32

        # pool = (likelihood pool, sampling pool); an int > 1 in the first
        # slot starts worker processes for a host likelihood
        # (sampler.py:283-298); a GpuPool names the GPUs of the batch
        try:
            pools = list(pool)
        except TypeError:
            pools = [pool]
        for i, p in enumerate(pools):
            if p is None or isinstance(p, GpuPool) or (
                    isinstance(p, int) and p == 1):
                pools[i] = p if isinstance(p, GpuPool) else None
            elif i == 0 and self.device_likelihood:
                # a device likelihood is evaluated inside the cycle kernels:
                # worker processes for it would never be used
                if len(pools) == 1:
                    raise ValueError(
                        'A DeviceLikelihood runs on the GPU; a process pool '
                        'has no use for it. Pass pool=GpuPool(n) to shard '
                        'the proposal batch over n GPUs.')
                pools[i] = None
            elif i == 0 and isinstance(p, int):
                pools[i] = NautilusPool(p, likelihood=self.likelihood)
                self.likelihood = likelihood_worker
            else:
                pools[i] = NautilusPool(p)
        self.pool_l = None if isinstance(pools[0], GpuPool) else pools[0]
        self.pool_s = pools[-1]

        if n_batch is None:
            s = 1 if self.pool_l is None else self.pool_l.size
            n_batch = -(-100 // s) * s
        self.n_batch = n_batch

        self.rng = np.random.default_rng(seed)

        self.n_like = 0
        self.explored = False
        self.bounds = []
        self.points = []
        self.log_l = []
        self.blobs = None
        self.blobs_dtype = blobs_dtype
        self._discard_exploration = False
        self.shell_n = np.zeros(0, dtype=int)
        self.shell_n_sample = np.zeros(0, dtype=int)
        self.shell_n_eff = np.zeros(0, dtype=float)
        self.shell_log_l_min = np.zeros(0, dtype=float)
        self.shell_log_l = np.zeros(0, dtype=float)
        self.shell_log_v = np.zeros(0, dtype=float)
        self.shell_n_sample_exp = np.zeros(0, dtype=int)
        self.shell_end_exp = np.zeros(0, dtype=int)
        self.points_t = np.zeros((0, self.n_dim))
        self.shell_t = np.zeros(0, dtype=int)
        self.log_l_t = np.zeros(0)
        self.blobs_t = None
        self.filepath = None
        self._stack = None       # all bounds serialised on the device
        self._like_params = None

    # ------------------------------------------------------------------
    # scheduler (sampler.py:373-505)
    # ------------------------------------------------------------------
    def run(self, f_live=0.01, n_shell=1, n_eff=10000, n_like_max=np.inf,
            discard_exploration=False, timeout=np.inf, verbose=False):
        """Run until convergence; returns False if stopped by ``n_like_max``
        or ``timeout``."""
        t_start = time()
        if verbose:
            print('Starting the nautilus_b200 sampler...' if self.n_like == 0
                  else 'Resuming nautilus_b200 run...')
            self.print_status(header=True)

        if len(self.bounds) == 0:
            self.add_bound()
            self.n_update_iter = -self.n_live
            self.n_like_iter = 0

        def done():
            return bool(self.explored and np.all(self.shell_n >= n_shell) and
                        self.n_eff >= n_eff)

        success = done()
        while (self.n_like < n_like_max and time() - t_start < timeout and
               not success):
            if not self.explored:
                if ((self.n_update_iter >= self.n_update or
                     self.n_like_iter >= self.n_like_new_bound) and
                        np.sum(self.shell_n) > self.n_live):
                    self.add_bound(verbose=verbose)
                    self.n_update_iter = 0
                    self.n_like_iter = 0
                self.n_update_iter += self.add_samples(-1, verbose=verbose)
                self.n_like_iter += self.n_batch
                if self.f_live <= f_live:
                    self._finish_exploration(discard_exploration)
            elif np.any(self.shell_n < n_shell):
                self.add_samples(int(np.flatnonzero(
                    self.shell_n < n_shell)[0]), verbose=verbose)
            elif self.n_eff < n_eff:
                gain = (self.shell_log_l + self.shell_log_v -
                        0.5 * np.log(self.shell_n) -
                        0.5 * np.log(self.shell_n_eff))
                self.add_samples(int(np.argmax(gain)), verbose=verbose)
            success = done()

        if verbose:
            self.print_status('Finished' if success else 'Stopped')
        return success

    def _finish_exploration(self, discard_exploration):
        """Drop empty shells and freeze the exploration bookkeeping
        (sampler.py:455-480)."""
        for shell in np.flatnonzero(self.shell_n == 0)[::-1]:
            del self.bounds[shell], self.points[shell], self.log_l[shell]
            if self.blobs is not None:
                del self.blobs[shell]
            for key in ('shell_n', 'shell_n_sample', 'shell_n_eff',
                        'shell_log_l_min', 'shell_log_l', 'shell_log_v'):
                setattr(self, key, np.delete(getattr(self, key), shell))
            self._stack = None
        self.shell_n_sample_exp = np.copy(self.shell_n_sample)
        self.shell_end_exp = np.array([len(p) for p in self.points])
        self.explored = True
        self.discard_exploration = discard_exploration

    @property
    def discard_exploration(self):
        return self._discard_exploration

    @discard_exploration.setter
    def discard_exploration(self, value):
        if not isinstance(value, bool):
            raise ValueError("'discard_exploration' must be a bool.")
        self._discard_exploration = value
        for index in range(len(self.log_l)):
            self.update_shell_info(index)

    # ------------------------------------------------------------------
    # read-outs (sampler.py:541-730, 1147-1190)
    # ------------------------------------------------------------------
    def _starts(self):
        if self._discard_exploration and self.explored:
            return self.shell_end_exp
        return np.zeros(len(self.points), dtype=int)

    def posterior(self, return_as_dict=None, equal_weight=False,
                  equal_weight_boost=1.0, return_blobs=False):
        """Posterior sample: points, log weights, log likelihoods[, blobs]."""
        if return_as_dict is None:
            return_as_dict = bool(callable(self.prior) and self.pass_dict)
        start = self._starts()
        points = np.concatenate([p[s:] for p, s in zip(self.points, start)])
        log_l = np.concatenate([v[s:] for v, s in zip(self.log_l, start)])
        log_w = np.repeat(self.shell_log_v - np.log(np.maximum(
            self.shell_n, 1)), self.shell_n) + log_l
        blobs = None
        if return_blobs:
            if self.blobs is None:
                raise ValueError('No blobs have been calculated.')
            blobs = np.concatenate(
                [b[s:] for b, s in zip(self.blobs, start)])

        if equal_weight:
            expect = np.exp(log_w - np.amax(log_w)) * equal_weight_boost
            whole = np.floor(expect)
            copies = whole.astype(int) + (
                self.rng.random(len(expect)) < expect - whole).astype(int)
            points = np.repeat(points, copies, axis=0)
            log_l = np.repeat(log_l, copies, axis=0)
            log_w = np.zeros(int(np.sum(copies)))
            if return_blobs:
                blobs = np.repeat(blobs, copies, axis=0)

        if callable(self.prior):
            transform = self.prior
        elif return_as_dict:
            transform = self.prior.unit_to_dictionary
        else:
            transform = self.prior.unit_to_physical
        if not self.vectorized and callable(self.prior):
            points = np.array(list(map(transform, points)))
        else:
            points = transform(points)
        if not return_as_dict and callable(self.prior) and self.pass_dict:
            raise ValueError('Cannot return points as numpy array. The prior '
                             'function only returns dictionaries.')

        log_w = log_w - logsumexp(log_w)
        if return_blobs:
            return points, log_w, log_l, blobs
        return points, log_w, log_l

    @property
    def n_eff(self):
        """Total effective sample size (sampler.py:650-665)."""
        if np.all(self.shell_n_eff == 0):
            return 0
        log_z = self.shell_log_l + self.shell_log_v
        used = self.shell_n_eff > 0
        w = np.exp(log_z - np.nanmax(log_z))[used]
        return np.sum(w)**2 / np.sum(w**2 / self.shell_n_eff[used])

    @property
    def log_z(self):
        """log evidence (sampler.py:681-694)."""
        if np.sum(self.shell_n) == 0:
            return None
        ok = ~np.isnan(self.shell_log_l)
        return logsumexp(self.shell_log_l[ok] + self.shell_log_v[ok])

    @property
    def eta(self):
        """Asymptotic sampling efficiency (sampler.py:709-730)."""
        ok = ~np.isnan(self.shell_log_l)
        log_z = (self.shell_log_l + self.shell_log_v)[ok]
        eff = (self.shell_n_eff / self.shell_n)[ok]
        return np.exp(2 * logsumexp(log_z) -
                      2 * logsumexp(log_z - 0.5 * np.log(eff)))

    def _live_weights(self):
        log_v = np.repeat(self.shell_log_v - np.log(np.maximum(
            self.shell_n, 1)), self.shell_n)
        log_l = np.concatenate(self.log_l)
        order = np.argsort(log_l)[-self.n_live:]
        return log_v, log_l, order

    @property
    def f_live(self):
        """Evidence fraction of the live set (sampler.py:1146-1169)."""
        if self.explored:
            return None
        if np.sum(self.shell_n) == 0:
            return 1.0
        log_v, log_l, live = self._live_weights()
        log_w = log_v + log_l
        return np.exp(logsumexp(log_w[live]) - logsumexp(log_w))

    @property
    def log_v_live(self):
        """log volume of the live set (sampler.py:1171-1190)."""
        if len(self.bounds) == 0:
            return 1.0
        log_v, _, live = self._live_weights()
        return logsumexp(log_v[live])

    # ------------------------------------------------------------------
    # device plumbing
    # ------------------------------------------------------------------
    def _device_stack(self):
        if self._stack is None:
            self._stack = ops.DeviceStack([b.spec() for b in self.bounds],
                                          device=default_device())
        return self._stack

    def _contains(self, index, points_dev, mask=None):
        return self._device_stack().contains(index, points_dev, mask=mask,
                                             mode=self.mlp_mode)

    def shell_association(self, points, n_max=None):
        """Index of the last bound (< n_max) containing each point
        (sampler.py:1192-1221)."""
        if n_max is None:
            n_max = len(self.bounds)
        dev = default_device()
        pts = torch.as_tensor(np.ascontiguousarray(points), device=dev)
        shell = torch.full((len(points), ), -1, dtype=torch.int64, device=dev)
        for i in range(n_max - 1, -1, -1):
            undecided = shell < 0
            if not bool(undecided.any()):
                break
            inside = self._contains(i, pts, mask=undecided)
            shell[undecided & inside] = i
        return shell.cpu().numpy()

    def shell_bound_occupation(self, fractional=True):
        """m[i, j] = how many points of shell i lie in bound j
        (sampler.py:1223-1251)."""
        n = len(self.bounds)
        m = np.zeros((n, n), dtype=int)
        dev = default_device()
        for i, points in enumerate(self.points):
            if len(points) == 0:
                continue
            pts = torch.as_tensor(np.ascontiguousarray(points), device=dev)
            for j in range(n):
                m[i, j] = int(self._contains(j, pts).sum().item())
        if fractional:
            # every stored point of shell i lies in bound i: the diagonal
            # counts the points (sampler.py:1248-1250)
            m = m / np.maximum(np.diag(m), 1)[:, np.newaxis]
        return m

    # ------------------------------------------------------------------
    # the cycle (sampler.py:751-943, 1093-1144)
    # ------------------------------------------------------------------
    def sample_shell(self, index, shell_t=None):
        """Exactly ``n_batch`` points uniform in shell ``index``
        (sampler.py:751-830).  Returns (points, n_bound[, idx_t])."""
        if shell_t is not None and index not in (-1, len(self.bounds) - 1):
            raise ValueError("'shell_t' must be empty list if not sampling "
                             "from the last bound/shell.")
        index = index % len(self.bounds)
        bound = self.bounds[index]
        n_bound, n_have = 0, 0
        idx_t = np.zeros(0, dtype=int)
        kept = []
        while n_have < self.n_batch:
            want = self.n_batch - n_have
            pts = (bound.sample(want, as_numpy=False, pool=self.pool_s)
                   if isinstance(self.pool_s, GpuPool)
                   else bound.sample(want, as_numpy=False))
            n_bound += want
            # drop points that belong to a later shell (every later bound is
            # consulted, as in the reference)
            alive = torch.ones(want, dtype=torch.bool, device=pts.device)
            for later in range(index + 1, len(self.bounds)):
                alive &= ~self._contains(later, pts, mask=alive)
            pts = pts[alive].cpu().numpy()

            replace = np.zeros(len(pts), dtype=bool)
            if shell_t is not None and len(shell_t) > 0 and len(pts) > 0:
                shell_p = self.shell_association(pts,
                                                 n_max=len(self.bounds) - 1)
                for shell in range(len(self.bounds) - 1):
                    donors = np.flatnonzero(shell_t == shell)
                    fresh = np.flatnonzero(shell_p == shell)
                    n = min(len(donors), len(fresh))
                    if n > 0:
                        idx_t = np.append(idx_t, self.rng.choice(
                            donors, size=n, replace=False))
                        shell_t[idx_t] = -1
                        replace[self.rng.choice(fresh, size=n,
                                                replace=False)] = True
            pts = pts[~replace]
            if len(pts) > 0:
                kept.append(pts)
                n_have += len(pts)
        points = np.concatenate(kept)
        if shell_t is None:
            return points, n_bound
        return points, n_bound, idx_t

    def evaluate_likelihood(self, points):
        """log L (and blobs) of unit-cube ``points`` (sampler.py:832-908)."""
        if self.device_likelihood:
            dev = default_device()
            if self._like_params is None:
                self._like_params = self.likelihood.device_params(dev)
            pts = torch.as_tensor(np.ascontiguousarray(points), device=dev)
            log_l = ops.loglike(pts, self.likelihood.like_id,
                                self._like_params).cpu().numpy()
            self.n_like += len(log_l)
            return log_l, None

        if callable(self.prior):
            transform = self.prior
        elif self.pass_dict:
            transform = self.prior.unit_to_dictionary
        else:
            transform = self.prior.unit_to_physical
        if self.vectorized:
            chunks = 1 if self.pool_l is None else self.pool_l.size
            args = [transform(c) for c in np.array_split(points, chunks)]
        else:
            args = [transform(p) for p in np.copy(points)]
        mapper = map if self.pool_l is None else self.pool_l.map
        result = list(mapper(self.likelihood, args))

        blobs = None
        if isinstance(result[0], tuple):
            blobs = [r[1:] for r in result]
            result = [r[0] for r in result]
        log_l = (np.concatenate(result) if self.vectorized
                 else np.array(result))
        if blobs is not None:
            join = np.concatenate if self.vectorized else np.array
            cols = [join([row[c] for row in blobs])
                    for c in range(len(blobs[0]))]
            if self.blobs_dtype is None:
                if len(cols) > 1:
                    self.blobs_dtype = [('blob_{}'.format(i), c.dtype)
                                        for i, c in enumerate(cols)]
                else:
                    self.blobs_dtype = cols[0].dtype
            blobs = np.squeeze(np.array(list(zip(*cols)),
                                        dtype=self.blobs_dtype))
        self.n_like += len(log_l)
        return log_l, blobs

    def update_shell_info(self, index):
        """Volume, mean likelihood and ESS of one shell from its stored log_l
        (sampler.py:910-943); the sums run on the GPU (nb200_stats)."""
        n_sample = self.shell_n_sample[index]
        start = 0
        if self._discard_exploration and self.explored:
            start = self.shell_end_exp[index]
            n_sample = n_sample - self.shell_n_sample_exp[index]
        log_l = self.log_l[index][start:]
        n = len(log_l)
        self.shell_n[index] = n
        if n == 0:
            self.shell_log_v[index] = -np.inf
            self.shell_log_l[index] = np.nan
            self.shell_n_eff[index] = 0
            return
        self.shell_log_v[index] = self.bounds[index].log_v + np.log(
            n / n_sample)
        lse, _ = ops.stats(torch.as_tensor(np.ascontiguousarray(log_l),
                                           device=default_device()))
        m, s1, s2 = lse.cpu().numpy()[:3]
        if s1 > 0:
            self.shell_log_l[index] = m + np.log(s1) - np.log(n)
            self.shell_n_eff[index] = s1 * s1 / s2
        else:                         # every log_l is -inf
            self.shell_log_l[index] = -np.inf
            self.shell_n_eff[index] = n

    def add_samples(self, shell, verbose=False):
        """One batch of ``n_batch`` likelihood evaluations in ``shell``
        (sampler.py:1093-1144).  Returns how many reach the shell's likelihood
        threshold."""
        if verbose:
            self.print_status('Sampling', end='\r')
        if shell == -1 and len(self.shell_t) > 0:
            points, n_bound, idx_t = self.sample_shell(-1, self.shell_t)
            assert len(points) + len(idx_t) == n_bound
            if len(idx_t) > 0:
                self.points[-1] = np.concatenate(
                    (self.points[-1], self.points_t[idx_t]))
                self.log_l[-1] = np.concatenate(
                    (self.log_l[-1], self.log_l_t[idx_t]))
                if self.blobs is not None:
                    self.blobs[-1] = np.concatenate(
                        (self.blobs[-1], self.blobs_t[idx_t]))
        else:
            points, n_bound = self.sample_shell(shell)
        if verbose:
            self.print_status('Computing', end='\r')

        self.shell_n_sample[shell] += n_bound
        log_l, blobs = self.evaluate_likelihood(points)
        self.points[shell] = np.append(self.points[shell], points, axis=0)
        self.log_l[shell] = np.append(self.log_l[shell], log_l, axis=0)
        if blobs is not None:
            if self.blobs is None:
                self.blobs = [blobs]
            else:
                self.blobs[shell] = np.append(self.blobs[shell], blobs, axis=0)
        self.update_shell_info(shell)
        return int(np.sum(log_l >= self.shell_log_l_min[shell]))

    # ------------------------------------------------------------------
    # new bounds (sampler.py:982-1091)
    # ------------------------------------------------------------------
    def add_bound(self, verbose=False):
        """Build the next bound from the live set; keep it only if smaller."""
        if len(self.bounds) == 0:
            log_l_min = -np.inf
            new_bound = UnitCube.compute(self.n_dim, rng=self.rng)
        else:
            if verbose:
                self.print_status('Bounding', end='\r')
            log_l = np.concatenate(self.log_l)
            order = np.argsort(log_l)
            points = np.concatenate(self.points)[order]
            log_l = log_l[order]
            log_l_min = log_l[-self.n_live]
            # step over a likelihood plateau if enough points lie above it
            above = log_l > log_l_min
            if (np.sum(log_l == log_l_min) > 1 and
                    np.sum(above) >= self.n_points_min):
                log_l_min = np.amin(log_l[above])
            new_bound = None
            if not np.all(log_l >= log_l_min):
                cand = NautilusBound.compute(
                    points, log_l, log_l_min, self.log_v_live,
                    enlarge_per_dim=self.enlarge_per_dim,
                    n_points_min=self.n_points_min,
                    split_threshold=self.split_threshold, periodic=None,
                    n_networks=self.n_networks,
                    neural_network_kwargs=self.neural_network_kwargs,
                    pool=None, rng=self.rng, mode=self.mlp_mode)
                cand.sample(1000, return_points=False)
                if cand.log_v < self.bounds[-1].log_v:
                    new_bound = cand
            if new_bound is None:
                self.shell_log_l_min[-1] = log_l_min
                return False

        self.bounds.append(new_bound)
        self._stack = None
        self.shell_n = np.append(self.shell_n, 0)
        self.shell_n_sample = np.append(self.shell_n_sample, 0)
        self.shell_n_eff = np.append(self.shell_n_eff, 0)
        self.shell_log_l = np.append(self.shell_log_l, np.nan)
        self.shell_log_v = np.append(self.shell_log_v, np.nan)
        self.shell_log_l_min = np.append(self.shell_log_l_min, log_l_min)
        self.points.append(np.zeros((0, self.n_dim)))
        self.log_l.append(np.zeros(0))
        if self.blobs is not None:
            self.blobs.append(np.zeros(self.blobs[-1][:0].shape,
                                       dtype=self.blobs_dtype))

        # earlier points inside the new bound become transfer candidates
        if len(self.bounds) > 1:
            moved = dict(shell=[], points=[], log_l=[], blobs=[])
            for shell in range(len(self.bounds) - 1):
                if len(self.points[shell]):
                    inside = self.bounds[-1].contains(
                        self.points[shell], mode=self.mlp_mode)
                else:
                    inside = np.zeros(0, dtype=bool)
                moved['shell'].append(np.repeat(shell, np.sum(inside)))
                moved['points'].append(self.points[shell][inside])
                moved['log_l'].append(self.log_l[shell][inside])
                self.points[shell] = self.points[shell][~inside]
                self.log_l[shell] = self.log_l[shell][~inside]
                if self.blobs is not None:
                    moved['blobs'].append(self.blobs[shell][inside])
                    self.blobs[shell] = self.blobs[shell][~inside]
                self.shell_n[shell] -= np.sum(inside)
                self.update_shell_info(shell)
            self.shell_t = np.concatenate(moved['shell'])
            self.points_t = np.concatenate(moved['points'])
            self.log_l_t = np.concatenate(moved['log_l'])
            if self.blobs is not None:
                self.blobs_t = np.concatenate(moved['blobs'])
        return True

    # ------------------------------------------------------------------
    def print_status(self, status='', header=False, end='\n'):
        """One status line (sampler.py:945-980)."""
        if header:
            cells = ['Status', 'Bounds', 'Ellipses', 'Networks', 'Calls',
                     'f_live', 'N_eff', 'log Z']
        else:
            last = self.bounds[-1] if len(self.bounds) > 1 else None
            values = [status, len(self.bounds),
                      last.n_ell if last else 0, last.n_net if last else 0,
                      self.n_like, self.f_live, self.n_eff, self.log_z]
            fmts = ['{}', '{:d}', '{:d}', '{:d}', '{:d}', '{:.4f}', '{:.0f}',
                    '{:+.2f}']
            cells = ['N/A' if v is None else f.format(v)
                     for v, f in zip(values, fmts)]
        widths = [9, 6, 8, 8, 8, 6, 5, 7]
        line = ' | '.join('{:<{}}'.format(c, w)
                          for c, w in zip(cells, widths))
        width = get_terminal_size((80, 24)).columns
        print(line.ljust(width)[:width], end=end, flush=True)
