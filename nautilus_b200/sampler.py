"""Importance nested sampling driver.

Drop-in for ``nautilus.Sampler`` (nautilus/sampler.py): same constructor and
``run`` / ``posterior`` signatures, same attributes (``bounds``, ``points``,
``log_l``, ``shell_*``), same control flow between shells -- with the
per-shell proposal -> neural filter -> exclusion -> likelihood ->
importance-weight cycle executed by the CUDA kernels behind
``include/nautilus_b200.h``.

Where things live.  The control plane (this file) is host Python like the
reference's.  Every evaluated point stays ON THE DEVICE in one arena (rows,
``log_l`` and a shell tag per point); ``sampler.points`` / ``sampler.log_l``
are lazy host views for user code.  With a ``DeviceLikelihood`` one step of
``add_samples`` is ONE ``nb200_cycle`` call per raw batch (proposal, neural
filter, exclusion by all later bounds, likelihood, log-sum-exp / counters) and
ONE 96-byte device->host read; shell sums are merged incrementally
(``update_shell_info`` never re-reads a shell's ``log_l`` from the host); the
live set (``f_live``, ``log_v_live``, the threshold of a new bound) and the
transfer of old points into a new bound are computed on the device.

Differences that are deliberate and documented in DESIGN.md:
  * random numbers come from Philox streams seeded by ``seed`` (the reference
    threads one PCG64 generator through everything), so runs are reproducible
    but not draw-for-draw identical to the reference;
  * ``likelihood`` may be a ``nautilus_b200.likelihoods.DeviceLikelihood``;
    it is then evaluated on the GPU inside the cycle.  On that path a step
    consumes WHOLE raw batches: ``n_batch`` is the least number of likelihood
    evaluations of a step, not the exact number (the bookkeeping --
    ``shell_n``, ``shell_n_sample``, the bound counters -- is exact);
  * checkpoints (``filepath=`` / ``resume=``, ``write``,
    ``write_shell_update``) keep the reference's layout; they go to HDF5 when
    h5py is installed and to the built-in '.npz' container (``_store.py``)
    otherwise.  A batch takes well under a millisecond here, so ``run``
    rewrites the file at most every ``checkpoint_interval`` seconds (30 by
    default; 0 = after every step, like the reference) and when it returns,
    not after every batch;
  * periodic parameters are out of scope and raise.
"""

from functools import partial
from pathlib import Path
from shutil import get_terminal_size
from time import time

import numpy as np
import torch
from scipy.special import logsumexp

from . import ops
from ._store import check_suffix, open_store, require_backend
from .bounds import NautilusBound, UnitCube
from ._device import default_device
from .likelihoods import DeviceLikelihood
from .pool import GpuPool, NautilusPool, likelihood_worker, merge_lse

TAG_DROPPED = -1          # arena tag of a point that belongs to no shell


class _Arena:
    """Every evaluated point, on the device: rows f64[n, d], ``log_l`` f64[n]
    and an int32 tag per point -- the index of its shell, ``TAG_DROPPED``, or
    ``-(2 + s)`` for a transfer candidate that came out of shell s
    (sampler.py:1059-1089).  Positions are chronological; nothing is ever
    moved, shells are told apart by the tag."""

    def __init__(self, n_dim, device, capacity=1 << 15):
        self.n_dim = n_dim
        self.device = device
        self.n = 0
        self.version = 0
        self.points = torch.empty((capacity, n_dim), dtype=torch.float64,
                                  device=device)
        self.log_l = torch.empty(capacity, dtype=torch.float64, device=device)
        self.tag = torch.empty(capacity, dtype=torch.int32, device=device)

    def reserve(self, extra):
        need = self.n + int(extra)
        cap = self.points.shape[0]
        if need <= cap:
            return
        while cap < need:
            cap *= 2
        for name in ('points', 'log_l', 'tag'):
            old = getattr(self, name)
            new = torch.empty((cap, ) + tuple(old.shape[1:]), dtype=old.dtype,
                              device=self.device)
            new[:self.n] = old[:self.n]
            setattr(self, name, new)

    def commit(self, k, shell):
        """k rows were written at [n, n + k): tag them, make them count."""
        self.tag[self.n:self.n + k] = shell
        self.n += int(k)
        self.version += 1

    def append(self, points, log_l, shell):
        points = torch.as_tensor(np.ascontiguousarray(points),
                                 device=self.device)
        k = points.shape[0]
        self.reserve(k)
        self.points[self.n:self.n + k] = points
        self.log_l[self.n:self.n + k] = torch.as_tensor(
            np.ascontiguousarray(log_l, dtype=np.float64), device=self.device)
        self.commit(k, shell)

    def view(self):
        return (self.points[:self.n], self.log_l[:self.n], self.tag[:self.n])


class Sampler:
    """Importance nested sampler (nautilus/sampler.py:21-1377)."""

    def __init__(self, prior, likelihood, n_dim=None, n_live=2000,
                 n_update=None, enlarge_per_dim=1.1, n_points_min=None,
                 split_threshold=100, periodic=None, n_networks=4,
                 neural_network_kwargs={}, prior_args=[], prior_kwargs={},
                 likelihood_args=[], likelihood_kwargs={}, n_batch=None,
                 n_like_new_bound=None, vectorized=False, pass_dict=None,
                 pool=None, seed=None, blobs_dtype=None, filepath=None,
                 resume=True, emulator_arith='auto', device_cycle=True,
                 checkpoint_interval=30.0):
        if filepath is not None:
            require_backend(filepath)
        if periodic is not None:
            raise NotImplementedError(
                'periodic parameters are outside the scope of nautilus_b200.')

        self.device_likelihood = isinstance(likelihood, DeviceLikelihood)
        # device_cycle=False keeps a DeviceLikelihood on the staged path of a
        # host likelihood (exactly n_batch evaluations per step, one op at a
        # time); tests use it as the reference semantics
        self.device_cycle = bool(device_cycle) and self.device_likelihood
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
        # 'auto': tensor cores (fp16 operands, else tf32) whenever the emulator
        # architecture fits the tcgen05 kernels, fp64 otherwise; 'f64' is the
        # bit-parity mode
        self.emulator_arith = emulator_arith
        self.mlp_mode = {'f64': ops.MLP_F64, 'tf32': ops.MLP_TF32,
                         'f16': ops.MLP_F16,
                         'auto': ops.MLP_F64}[emulator_arith]
        if emulator_arith == 'auto':
            from ._pack import pack_tc, pack_tc16
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
                # fp16 operands (same 10 mantissa bits, half the bytes) when
                # the weights fit shared memory in that form
                if pack_tc16(probe, 0.0) is not None:
                    self.mlp_mode = ops.MLP_F16

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
            if self.device_cycle:
                # a step is whole raw GPU batches: ask for
                # enough points per step to fill one
                n_batch = max(n_batch, n_live // 2)
        self.n_batch = n_batch

        self.rng = np.random.default_rng(seed)

        self.n_like = 0
        self.explored = False
        self.bounds = []
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
        self.filepath = filepath
        self.checkpoint_interval = float(checkpoint_interval)
        self._t_checkpoint = -np.inf    # time of the last write of the file
        self._dirty = False             # state newer than the file
        self._arena = None          # created with the first bound
        self._blobs_all = None      # host, aligned with arena positions
        self._host_cache = (None, None)
        self._sums = []             # per shell: running (m, s1, s2) or None
        self._explore_end = 0       # arena position where exploration ended
        self._t_pos = np.zeros(0, dtype=np.int64)   # transfer candidates:
        self._t_shell = np.zeros(0, dtype=int)      # arena position, donor
        self._p_shell = {}          # per bound: (in-shell, raw) so far
        self._stack = None          # all bounds serialised on the device
        self._stack_len = 0
        self._like_params = None
        self._cycle_buf = None
        self.cycle_stats = dict(calls=0, raw=0, d2h_bytes=0)
        if resume and filepath is not None and Path(filepath).exists():
            self._resume(filepath)

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
                    self._checkpoint()
                n_like_before = self.n_like
                self.n_update_iter += self.add_samples(-1, verbose=verbose)
                self.n_like_iter += self.n_like - n_like_before
                self._checkpoint()
                if self.f_live <= f_live:
                    self._finish_exploration(discard_exploration)
                    self._checkpoint()
            elif np.any(self.shell_n < n_shell):
                self.add_samples(int(np.flatnonzero(
                    self.shell_n < n_shell)[0]), verbose=verbose)
                self._checkpoint()
            elif self.n_eff < n_eff:
                gain = (self.shell_log_l + self.shell_log_v -
                        0.5 * np.log(self.shell_n) -
                        0.5 * np.log(self.shell_n_eff))
                self.add_samples(int(np.argmax(gain)), verbose=verbose)
                self._checkpoint()
            success = done()

        # whatever happened since the last write (the reference's file is
        # current after every batch, sampler.py:449-494; here a write moves
        # every stored point to the host and a whole run takes seconds, so
        # the file is at most `checkpoint_interval` seconds old while the run
        # lasts and current when it returns)
        if self._dirty:
            self._checkpoint(force=True)
        if verbose:
            self.print_status('Finished' if success else 'Stopped')
        return success

    def _finish_exploration(self, discard_exploration):
        """Drop empty shells and freeze the exploration bookkeeping
        (sampler.py:455-480)."""
        _, _, tag = self._arena.view()
        # whatever is still waiting for a transfer belongs to no shell
        tag[tag <= -2] = TAG_DROPPED
        self._t_pos = np.zeros(0, dtype=np.int64)
        self._t_shell = np.zeros(0, dtype=int)
        empty = np.flatnonzero(self.shell_n == 0)
        if len(empty):
            keep = np.setdiff1d(np.arange(len(self.bounds)), empty)
            remap = np.full(len(self.bounds), TAG_DROPPED, dtype=np.int32)
            remap[keep] = np.arange(len(keep), dtype=np.int32)
            lut = torch.as_tensor(remap, device=tag.device)
            inside = tag >= 0
            tag[inside] = lut[tag[inside].long()]
            self.bounds = [self.bounds[i] for i in keep]
            self._sums = [self._sums[i] for i in keep]
            for key in ('shell_n', 'shell_n_sample', 'shell_n_eff',
                        'shell_log_l_min', 'shell_log_l', 'shell_log_v'):
                setattr(self, key, getattr(self, key)[keep])
            self._stack = None
        self._arena.version += 1
        self._p_shell = {}
        self.shell_n_sample_exp = np.copy(self.shell_n_sample)
        self.shell_end_exp = np.copy(self.shell_n)
        self._explore_end = self._arena.n
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
        for index in range(len(self.bounds)):
            self._sums[index] = None
            self.update_shell_info(index)

    # ------------------------------------------------------------------
    # stored points: device arena, lazy host views
    # ------------------------------------------------------------------
    def _discarding(self):
        return bool(self._discard_exploration and self.explored)

    def _start(self):
        """First arena position that counts (discard_exploration drops what
        was drawn while exploring, sampler.py:917-923)."""
        return self._explore_end if self._discarding() else 0

    def _host(self):
        """(points, log_l, tag) of the arena as NumPy arrays (cached until
        the arena changes)."""
        if self._arena is None:
            return (np.zeros((0, self.n_dim)), np.zeros(0),
                    np.zeros(0, dtype=np.int32))
        version, cached = self._host_cache
        if version != self._arena.version:
            pts, ll, tag = self._arena.view()
            cached = (pts.cpu().numpy(), ll.cpu().numpy(), tag.cpu().numpy())
            self._host_cache = (self._arena.version, cached)
        return cached

    def _by_shell(self, values, tag):
        return [values[tag == i] for i in range(len(self.bounds))]

    @property
    def points(self):
        """List (one entry per shell) of the stored points, as in the
        reference; materialised from the device arena on demand."""
        pts, _, tag = self._host()
        return self._by_shell(pts, tag)

    @property
    def log_l(self):
        _, ll, tag = self._host()
        return self._by_shell(ll, tag)

    @property
    def blobs(self):
        if self._blobs_all is None:
            return None
        _, _, tag = self._host()
        return self._by_shell(self._blobs_all[:len(tag)], tag)

    @property
    def points_t(self):
        return self._host()[0][self._t_pos]

    @property
    def log_l_t(self):
        return self._host()[1][self._t_pos]

    @property
    def shell_t(self):
        return self._t_shell

    # ------------------------------------------------------------------
    # read-outs (sampler.py:541-730, 1147-1190)
    # ------------------------------------------------------------------
    def posterior(self, return_as_dict=None, equal_weight=False,
                  equal_weight_boost=1.0, return_blobs=False):
        """Posterior sample: points, log weights, log likelihoods[, blobs]."""
        if return_as_dict is None:
            return_as_dict = bool(callable(self.prior) and self.pass_dict)
        pts, ll, tag = self._host()
        use = np.flatnonzero((tag >= 0) &
                             (np.arange(len(tag)) >= self._start()))
        # shell by shell, chronological inside a shell (the reference's order)
        use = use[np.argsort(tag[use], kind='stable')]
        points, log_l = pts[use], ll[use]
        log_w = np.repeat(self.shell_log_v - np.log(np.maximum(
            self.shell_n, 1)), self.shell_n) + log_l
        blobs = None
        if return_blobs:
            if self._blobs_all is None:
                raise ValueError('No blobs have been calculated.')
            blobs = self._blobs_all[use]

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

    def _live_set(self):
        """Device side of ``f_live`` / ``log_v_live`` / the threshold of a new
        bound (sampler.py:1007-1009, 1146-1190): over the m stored points
        that belong to a shell, CUDA tensors (position int64[m], log volume
        weight f64[m], log_l f64[m], live int64[<= n_live]) with ``live`` the
        indices of the n_live points of highest likelihood (device top-k)."""
        _, ll, tag = self._arena.view()
        pos = torch.nonzero(tag >= 0).squeeze(1)
        per_shell = torch.as_tensor(
            self.shell_log_v - np.log(np.maximum(self.shell_n, 1)),
            device=ll.device)
        log_l = ll[pos]
        log_v = per_shell[tag[pos].long()]
        live = ops.top_k(log_l, min(self.n_live, log_l.numel()))
        return pos, log_v, log_l, live

    @property
    def f_live(self):
        """Evidence fraction of the live set (sampler.py:1146-1169)."""
        if self.explored:
            return None
        if np.sum(self.shell_n) == 0:
            return 1.0
        _, log_v, log_l, live = self._live_set()
        log_w = log_v + log_l
        return float(torch.exp(torch.logsumexp(log_w[live], 0) -
                               torch.logsumexp(log_w, 0)).item())

    @property
    def log_v_live(self):
        """log volume of the live set (sampler.py:1171-1190)."""
        if len(self.bounds) == 0:
            return 1.0
        _, log_v, _, live = self._live_set()
        return float(torch.logsumexp(log_v[live], 0).item())

    # ------------------------------------------------------------------
    # device plumbing
    # ------------------------------------------------------------------
    def _device_stack(self):
        """All bounds serialised on the device; a new bound is APPENDED (only
        its own parameters are packed and uploaded)."""
        if self._stack is None:
            self._stack = ops.DeviceStack([b.spec() for b in self.bounds],
                                          device=default_device())
            self._stack_len = len(self.bounds)
        while self._stack_len < len(self.bounds):
            self._stack.append(self.bounds[self._stack_len].spec())
            self._stack_len += 1
        return self._stack

    def _contains(self, index, points_dev, mask=None):
        return self._device_stack().contains(index, points_dev, mask=mask,
                                             mode=self.mlp_mode)

    def _association(self, pts, n_max):
        """Device tensor: index of the last bound (< n_max) containing each
        row of the CUDA tensor ``pts`` (sampler.py:1213-1221)."""
        shell = torch.full((pts.shape[0], ), -1, dtype=torch.int64,
                           device=pts.device)
        for i in range(n_max - 1, -1, -1):
            undecided = shell < 0
            if not bool(undecided.any()):
                break
            inside = self._contains(i, pts, mask=undecided)
            shell[undecided & inside] = i
        return shell

    def shell_association(self, points, n_max=None):
        """Index of the last bound (< n_max) containing each point
        (sampler.py:1192-1221)."""
        if n_max is None:
            n_max = len(self.bounds)
        pts = torch.as_tensor(np.ascontiguousarray(points),
                              device=default_device())
        return self._association(pts, n_max).cpu().numpy()

    def shell_bound_occupation(self, fractional=True):
        """m[i, j] = how many points of shell i lie in bound j
        (sampler.py:1223-1251)."""
        n = len(self.bounds)
        m = np.zeros((n, n), dtype=int)
        pts, _, tag = self._arena.view()
        pos = torch.nonzero(tag >= 0).squeeze(1)
        if pos.numel():
            rows = pts[pos].contiguous()
            shell = tag[pos].long()
            for j in range(n):
                inside = self._contains(j, rows)
                m[:, j] = torch.bincount(shell[inside],
                                         minlength=n).cpu().numpy()
        if fractional:
            # every stored point of shell i lies in bound i: the diagonal
            # counts the points (sampler.py:1248-1250)
            m = m / np.maximum(np.diag(m), 1)[:, np.newaxis]
        return m

    # ------------------------------------------------------------------
    # the cycle on the device (sampler.py:751-943, 1093-1144 in one call)
    # ------------------------------------------------------------------
    def _raw_batch_size(self, index, want):
        """Raw proposals to draw so that about ``want`` end up in the shell,
        from the shell's own acceptance so far."""
        got, raw = self._p_shell.get(index, (0, 0))
        if raw > 0:
            p = max(got, 1) / raw
        else:
            bound = self.bounds[index]
            p = 1.0
            if isinstance(bound, NautilusBound) and bound.n_sample > 0 and \
                    bound.outer_bound.n_sample > 0:
                p = ((1 - bound.outer_bound.n_reject /
                      bound.outer_bound.n_sample) *
                     (1 - bound.n_reject / bound.n_sample))
            if index + 1 < len(self.bounds):
                p *= 0.5
        n_raw = int(1.15 * want / max(p, 1e-5)) + 256
        return int(min(max(n_raw, 4096), 1 << 22))

    def _cycle_buffers(self, n_raw):
        buf = self._cycle_buf
        if buf is None or buf['code'].numel() < n_raw:
            dev = default_device()
            cap = max(n_raw, 1 << 16)
            buf = dict(
                points=torch.empty((cap, self.n_dim), dtype=torch.float64,
                                   device=dev),
                log_l=torch.empty(cap, dtype=torch.float64, device=dev),
                code=torch.empty(cap, dtype=torch.uint8, device=dev),
                lse=torch.empty(ops.N_LSE, dtype=torch.float64, device=dev),
                counters=torch.empty(ops.N_CNT, dtype=torch.int64,
                                     device=dev))
            self._cycle_buf = buf
        return buf

    def _device_args(self, pts):
        """What the likelihood is called with (sampler.py:856-862), for rows
        that live on the device."""
        if callable(self.prior):
            return self.prior(pts)
        if self.pass_dict:
            return self.prior.unit_to_dictionary_device(pts)
        return self.prior.unit_to_physical_device(pts)

    def _run_cycle_plugin(self, index, n_raw, offset, out_points, out_log_l):
        """The cycle with a plug-in (torch-callable) likelihood: nb200_cycle
        without a built-in likelihood, the callable on the compacted rows,
        nb200_stats.  One extra 8-byte read (the row count)."""
        stack = self._device_stack()
        bound = self.bounds[index]
        buf = self._cycle_buffers(n_raw)
        out = stack.cycle(
            index, n_raw, later=(index + 1, len(self.bounds) - index - 1),
            seed=bound.stream.seed, offset=offset,
            stream_id=bound.stream.stream_id, like_id=-1, mode=self.mlp_mode,
            out=buf)
        _, _, n_out = stack.compact(out['points'][:n_raw], None,
                                    out['code'][:n_raw],
                                    out_points=out_points)
        k = int(n_out.item())
        counters = out['counters'].clone()
        if k == 0:
            lse = torch.tensor([-np.inf, 0.0, 0.0, 0.0], dtype=torch.float64,
                               device=stack.device)
            return counters, lse
        log_l = self.likelihood.torch_call(
            self._device_args(out_points[:k]))
        if log_l.numel() != k:
            raise ValueError('the likelihood returned {} values for {} '
                             'points'.format(log_l.numel(), k))
        out_log_l[:k] = log_l
        lse, cnt = ops.stats(log_l, log_l_min=float(
            self.shell_log_l_min[index]))
        counters[ops.CNT_UPDATE] = cnt[ops.CNT_UPDATE]
        return counters, lse

    def _run_cycle(self, index, n_raw, offset, out_points, out_log_l):
        """ONE nb200_cycle call: raw proposals [offset, offset + n_raw) of
        bound ``index`` -> neural filter -> exclusion by every later bound ->
        likelihood -> sums; the in-shell rows are compacted straight into
        ``out_points`` / ``out_log_l`` (views of the arena).  Returns the
        CUDA tensors (counters i64[8], lse f64[4]); nothing is synchronised.
        """
        if self.likelihood.like_id < 0:
            return self._run_cycle_plugin(index, n_raw, offset, out_points,
                                          out_log_l)
        stack = self._device_stack()
        if self._like_params is None:
            self._like_params = self.likelihood.device_params(stack.device)
        bound = self.bounds[index]
        buf = self._cycle_buffers(n_raw)
        out = stack.cycle(
            index, n_raw, later=(index + 1, len(self.bounds) - index - 1),
            seed=bound.stream.seed, offset=offset,
            stream_id=bound.stream.stream_id,
            like_id=self.likelihood.like_id, like_params=self._like_params,
            log_l_min=float(self.shell_log_l_min[index]), mode=self.mlp_mode,
            out=buf)
        stack.compact(out['points'][:n_raw], out['log_l'][:n_raw],
                      out['code'][:n_raw], out_points=out_points,
                      out_log_l=out_log_l)
        return out['counters'], out['lse']

    def _add_samples_device(self, index):
        """``add_samples`` with a device likelihood: whole raw batches
        through ``nb200_cycle`` until at least ``n_batch`` new points are in
        the shell.  Per raw batch the host reads 96 bytes."""
        bound = self.bounds[index]
        arena = self._arena
        last = index == len(self.bounds) - 1
        n_new, n_update = 0, 0
        while n_new < self.n_batch:
            n_raw = self._raw_batch_size(index, self.n_batch - n_new)
            offset = bound.stream.take(n_raw)
            arena.reserve(n_raw)
            a0 = arena.n
            counters, lse = self._run_cycle(
                index, n_raw, offset, arena.points[a0:a0 + n_raw],
                arena.log_l[a0:a0 + n_raw])
            # the one device->host read of the batch: 8 counters + the
            # log-sum-exp triple
            small = torch.cat([counters.double(), lse]).cpu().numpy()
            self.cycle_stats['calls'] += 1
            self.cycle_stats['raw'] += n_raw
            self.cycle_stats['d2h_bytes'] += small.nbytes
            cnt = small[:ops.N_CNT].round().astype(np.int64)
            triple = tuple(small[ops.N_CNT:ops.N_CNT + 3])
            k = int(cnt[ops.CNT_IN_SHELL])
            arena.commit(k, index)
            got, raw = self._p_shell.get(index, (0, 0))
            self._p_shell[index] = (got + k, raw + n_raw)
            # the bound's own counters (union.py:322-323, nautilus.py:221-222)
            if isinstance(bound, NautilusBound):
                rej_u = int(cnt[ops.CNT_CUBE_REJECT] +
                            cnt[ops.CNT_OVERLAP_REJECT])
                bound.outer_bound.n_sample += n_raw
                bound.outer_bound.n_reject += rej_u
                bound.n_sample += n_raw - rej_u
                bound.n_reject += int(cnt[ops.CNT_NN_REJECT])
            # draws of the bound this batch consumed (sampler.py:793, 1133)
            self.shell_n_sample[index] += k + int(cnt[ops.CNT_EXCLUDED])

            kept, upd = k, int(cnt[ops.CNT_UPDATE])
            if last and len(self._t_pos) > 0 and k > 0:
                kept, upd = self._swap_in_transfers(index, a0, k)
                self._sums[index] = None        # the set changed: re-reduce
            elif self._sums[index] is not None:
                self._sums[index] = merge_lse([self._sums[index], triple])
                self.shell_n[index] += k
            self.n_like += kept
            n_new += kept
            n_update += upd
            self.update_shell_info(index)
        return n_update

    def _swap_in_transfers(self, index, a0, k):
        """Exploration only (sampler.py:803-819): fresh draws that fall into
        an earlier shell s are replaced, one for one, by transfer candidates
        that came out of shell s.  Works on arena tags; returns (fresh points
        kept, how many of them reach the shell's threshold)."""
        arena = self._arena
        fresh = arena.points[a0:a0 + k]
        shell_p = self._association(fresh, len(self.bounds) - 1).cpu().numpy()
        drop, take = [], []
        for s in range(len(self.bounds) - 1):
            donors = np.flatnonzero(self._t_shell == s)
            cand = np.flatnonzero(shell_p == s)
            n = min(len(donors), len(cand))
            if n > 0:
                chosen = self.rng.choice(donors, size=n, replace=False)
                take.append(chosen)
                self._t_shell[chosen] = -1
                drop.append(self.rng.choice(cand, size=n, replace=False))
        if take:
            take = np.concatenate(take)
            drop = np.concatenate(drop)
            dev = arena.tag.device
            arena.tag[torch.as_tensor(self._t_pos[take], device=dev)] = index
            arena.tag[torch.as_tensor(a0 + drop, device=dev)] = TAG_DROPPED
            arena.version += 1
            still = self._t_shell >= 0
            self._t_pos = self._t_pos[still]
            self._t_shell = self._t_shell[still]
        fresh_in = arena.tag[a0:a0 + k] == index
        kept = int(fresh_in.sum().item())
        upd = int((fresh_in & (arena.log_l[a0:a0 + k] >=
                               float(self.shell_log_l_min[index]))
                   ).sum().item())
        return kept, upd

    # ------------------------------------------------------------------
    # the cycle with a host likelihood (sampler.py:751-908)
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
            pts = pts[alive].contiguous()

            replace = np.zeros(pts.shape[0], dtype=bool)
            if shell_t is not None and len(shell_t) > 0 and len(pts) > 0:
                shell_p = self._association(
                    pts, len(self.bounds) - 1).cpu().numpy()
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
            pts = pts.cpu().numpy()[~replace]
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
            pts = torch.as_tensor(np.ascontiguousarray(points), device=dev)
            if self.likelihood.like_id < 0:
                log_l = self.likelihood.torch_call(
                    self._device_args(pts)).cpu().numpy()
            else:
                if self._like_params is None:
                    self._like_params = self.likelihood.device_params(dev)
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

    def _add_samples_host(self, index):
        """``add_samples`` with a Python likelihood: exactly ``n_batch``
        evaluations (sampler.py:1093-1144); the new points are uploaded into
        the arena."""
        arena = self._arena
        last = index == len(self.bounds) - 1
        if last and len(self._t_pos) > 0:
            shell_t = self._t_shell.copy()
            points, n_bound, idx_t = self.sample_shell(index, shell_t)
            assert len(points) + len(idx_t) == n_bound
            if len(idx_t) > 0:
                arena.tag[torch.as_tensor(self._t_pos[idx_t],
                                          device=arena.device)] = index
                arena.version += 1
                still = np.ones(len(self._t_pos), dtype=bool)
                still[idx_t] = False
                self._t_pos = self._t_pos[still]
                self._t_shell = self._t_shell[still]
        else:
            points, n_bound = self.sample_shell(index)

        self.shell_n_sample[index] += n_bound
        log_l, blobs = self.evaluate_likelihood(points)
        if blobs is not None:
            blobs = np.atleast_1d(blobs)
            if self._blobs_all is None:
                self._blobs_all = np.zeros(arena.n, dtype=blobs.dtype)
            self._blobs_all = np.concatenate([self._blobs_all, blobs])
        arena.append(points, log_l, index)
        self._sums[index] = None
        self.update_shell_info(index)
        return int(np.sum(log_l >= self.shell_log_l_min[index]))

    # ------------------------------------------------------------------
    def update_shell_info(self, index):
        """Volume, mean likelihood and ESS of one shell (sampler.py:910-943).

        The log-sum-exp triple of a shell is the running merge of the triples
        ``nb200_cycle`` returns batch by batch (``self._sums``, with
        ``shell_n`` counted along); when the shell's membership changed in
        another way (transfers, discard_exploration, a host likelihood) it is
        re-reduced on the device from the arena (``nb200_stats``) -- the
        shell's ``log_l`` is never uploaded again."""
        n_sample = self.shell_n_sample[index]
        if self._discarding():
            n_sample = n_sample - self.shell_n_sample_exp[index]
        if self._sums[index] is None:
            _, ll, tag = self._arena.view()
            start = self._start()
            code = (tag[start:] == index).to(torch.uint8) * ops.CODE_IN_SHELL
            lse, cnt = ops.stats(ll[start:], code=code)
            small = torch.cat([cnt.double(), lse]).cpu().numpy()
            self.shell_n[index] = int(round(small[ops.CNT_IN_SHELL]))
            self._sums[index] = tuple(small[ops.N_CNT:ops.N_CNT + 3])
        n = int(self.shell_n[index])
        if n == 0:
            self.shell_log_v[index] = -np.inf
            self.shell_log_l[index] = np.nan
            self.shell_n_eff[index] = 0
            return
        self.shell_log_v[index] = self.bounds[index].log_v + np.log(
            n / n_sample)
        m, s1, s2 = self._sums[index]
        if s1 > 0:
            self.shell_log_l[index] = m + np.log(s1) - np.log(n)
            self.shell_n_eff[index] = s1 * s1 / s2
        else:                         # every log_l is -inf
            self.shell_log_l[index] = -np.inf
            self.shell_n_eff[index] = n

    def add_samples(self, shell, verbose=False):
        """One batch of likelihood evaluations in ``shell``
        (sampler.py:1093-1144).  Returns how many reach the shell's
        likelihood threshold."""
        if verbose:
            self.print_status('Sampling', end='\r')
        index = shell % len(self.bounds)
        if self.device_cycle and not isinstance(self.pool_s, GpuPool):
            return self._add_samples_device(index)
        return self._add_samples_host(index)

    # ------------------------------------------------------------------
    # new bounds (sampler.py:982-1091)
    # ------------------------------------------------------------------
    def add_bound(self, verbose=False):
        """Build the next bound from the live set; keep it only if smaller."""
        if len(self.bounds) == 0:
            log_l_min = -np.inf
            new_bound = UnitCube.compute(self.n_dim, rng=self.rng)
            self._arena = _Arena(self.n_dim, default_device())
        else:
            if verbose:
                self.print_status('Bounding', end='\r')
            # live threshold and plateau rule (sampler.py:1007-1016), on the
            # device: the n_live-th largest stored log_l
            pts, _, _ = self._arena.view()
            pos, log_v, log_l, live = self._live_set()
            log_l_min = float(log_l[live].min().item())
            above = log_l > log_l_min
            if int((log_l == log_l_min).sum().item()) > 1 and \
                    int(above.sum().item()) >= self.n_points_min:
                log_l_min = float(log_l[above].min().item())
            new_bound = None
            if not bool((log_l >= log_l_min).all().item()):
                # rows ordered by likelihood, as the reference hands them
                # over; they stay on the device (NautilusBound.compute pulls
                # the live points and the training subsets only)
                order = torch.argsort(log_l, stable=True)
                cand = NautilusBound.compute(
                    pts[pos[order]], log_l[order], log_l_min,
                    float(torch.logsumexp(log_v[live], 0).item()),
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
        self.shell_n = np.append(self.shell_n, 0)
        self.shell_n_sample = np.append(self.shell_n_sample, 0)
        self.shell_n_eff = np.append(self.shell_n_eff, 0)
        self.shell_log_l = np.append(self.shell_log_l, np.nan)
        self.shell_log_v = np.append(self.shell_log_v, np.nan)
        self.shell_log_l_min = np.append(self.shell_log_l_min, log_l_min)
        self._sums.append((-np.inf, 0.0, 0.0))

        # earlier points inside the new bound become transfer candidates
        # (sampler.py:1059-1089), found with ONE contains() over all stored
        # points on the device; candidates of the previous bound that were
        # never used belong to no shell any more
        if len(self.bounds) > 1:
            pts, _, tag = self._arena.view()
            tag[tag <= -2] = TAG_DROPPED
            pos = torch.nonzero(tag >= 0).squeeze(1)
            gone = np.zeros(len(self.bounds), dtype=int)
            self._t_pos = np.zeros(0, dtype=np.int64)
            self._t_shell = np.zeros(0, dtype=int)
            if pos.numel():
                inside = self._contains(len(self.bounds) - 1,
                                        pts[pos].contiguous())
                moved = pos[inside]
                donors = tag[moved].long()
                tag[moved] = (-2 - donors).to(torch.int32)
                gone = torch.bincount(
                    donors, minlength=len(self.bounds)).cpu().numpy()
                self._t_pos = moved.cpu().numpy()
                self._t_shell = donors.cpu().numpy().astype(int)
            self._arena.version += 1
            for shell in range(len(self.bounds) - 1):
                if gone[shell] > 0:
                    self._sums[shell] = None
                    self.update_shell_info(shell)
        return True

    # ------------------------------------------------------------------
    # checkpoints (sampler.py:329-371, 1253-1377)
    # ------------------------------------------------------------------
    _ATTRS_CONFIG = ['n_dim', 'n_live', 'n_update', 'n_like_new_bound',
                     'enlarge_per_dim', 'n_points_min', 'split_threshold',
                     'n_networks', 'n_batch', 'vectorized', 'pass_dict']
    _ATTRS_STATE = ['n_like', 'explored', '_discard_exploration', 'shell_n',
                    'shell_n_sample', 'shell_n_eff', 'shell_log_l_min',
                    'shell_log_l', 'shell_log_v', 'shell_n_sample_exp',
                    'shell_end_exp', 'n_update_iter', 'n_like_iter']

    def _checkpoint(self, force=False):
        """``run``'s hook: rewrite the file now (``force``) or if the last
        write is older than ``checkpoint_interval`` seconds."""
        if self.filepath is None:
            return
        self._dirty = True
        if force or time() - self._t_checkpoint >= self.checkpoint_interval:
            self.write(self.filepath, overwrite=True)

    def _write_rng(self, group):
        state = self.rng.bit_generator.state
        group.attrs['rng_state'] = str(state['state']['state'])
        group.attrs['rng_inc'] = str(state['state']['inc'])
        group.attrs['rng_has_uint32'] = state['has_uint32']
        group.attrs['rng_uinteger'] = state['uinteger']

    def _shell_arrays(self):
        """Per shell (points, log_l, blobs) and the transfer candidates, as
        the reference stores them."""
        pts, ll, tag = self._host()
        shells = []
        for i in range(len(self.bounds)):
            sel = tag == i
            shells.append((pts[sel], ll[sel], None if self._blobs_all is None
                           else self._blobs_all[:len(tag)][sel]))
        blobs_t = (None if self._blobs_all is None
                   else self._blobs_all[self._t_pos])
        return shells, (pts[self._t_pos], ll[self._t_pos], blobs_t)

    def write(self, filepath, overwrite=False):
        """Write the sampler to disk in the reference's layout
        (sampler.py:1253-1332): group 'sampler' with the settings and shell
        bookkeeping as attributes and 'points_<i>' / 'log_l_<i>'
        [/ 'blobs_<i>'] per shell plus the transfer candidates, and one group
        'bound_<i>' per bound.  '.h5' / '.hdf5' need h5py; '.npz' is the
        built-in container.

        Raises ValueError for another file extension and RuntimeError if the
        file exists and ``overwrite`` is False."""
        filepath = Path(filepath)
        check_suffix(filepath)
        if filepath.exists() and not overwrite:
            raise RuntimeError(
                'File {} already exists.'.format(str(filepath)))
        filepath.parent.mkdir(parents=True, exist_ok=True)
        # ('w': the '.npz' container replaces the old file atomically when it
        # is closed, HDF5 truncates)
        fstream = open_store(filepath, 'w')
        group = fstream.create_group('sampler')
        if len(self.bounds) > 0 and not hasattr(self, 'n_update_iter'):
            self.n_update_iter, self.n_like_iter = -self.n_live, 0
        for key in self._ATTRS_CONFIG + self._ATTRS_STATE:
            if hasattr(self, key):
                group.attrs[key] = getattr(self, key)
        for key, value in self.neural_network_kwargs.items():
            group.attrs['neural_network_{}'.format(key)] = value
        # not in the reference: where the exploration phase ended, per shell,
        # is 'shell_end_exp'; the emulator arithmetic of the bounds
        group.attrs['emulator_arith'] = self.emulator_arith
        # acceptance seen per shell so far (sizes the next raw batch)
        group.attrs['p_shell'] = np.array(
            [[i, got, raw] for i, (got, raw) in sorted(self._p_shell.items())],
            dtype=np.int64).reshape(-1, 3)

        shells, (points_t, log_l_t, blobs_t) = self._shell_arrays()
        for shell, (points, log_l, blobs) in enumerate(shells):
            group.create_dataset('points_{}'.format(shell), data=points,
                                 maxshape=(None, self.n_dim))
            group.create_dataset('log_l_{}'.format(shell), data=log_l,
                                 maxshape=(None, ))
            if blobs is not None:
                group.create_dataset(
                    'blobs_{}'.format(shell), data=blobs,
                    maxshape=(None, ) + tuple(blobs.shape[1:]))
        group.create_dataset('points_t', data=points_t,
                             maxshape=(None, self.n_dim))
        group.create_dataset('shell_t', data=self._t_shell,
                             maxshape=(None, ))
        group.create_dataset('log_l_t', data=log_l_t, maxshape=(None, ))
        if blobs_t is not None:
            group.create_dataset('blobs_t', data=blobs_t,
                                 maxshape=(None, ) + tuple(blobs_t.shape[1:]))

        for i, bound in enumerate(self.bounds):
            bound.write(fstream.create_group('bound_{}'.format(i)))
        self._write_rng(group)
        fstream.close()
        self._t_checkpoint = time()
        self._dirty = False

    def write_shell_update(self, filepath, shell):
        """Update the data of one shell in an existing file
        (sampler.py:1334-1377)."""
        if shell < 0:
            shell = len(self.bounds) + shell
        fstream = open_store(Path(filepath), 'r+')
        group = fstream['sampler']
        for key in ['n_like', 'shell_n', 'shell_n_sample', 'shell_n_eff',
                    'shell_log_l_min', 'shell_log_l', 'shell_log_v',
                    'n_update_iter', 'n_like_iter']:
            group.attrs[key] = getattr(self, key)

        def store(name, data):
            if data is None:
                return
            if name not in group:
                group.create_dataset(
                    name, data=data,
                    maxshape=(None, ) + tuple(data.shape[1:]))
                return
            group[name].resize(data.shape)
            group[name][...] = data

        shells, (points_t, log_l_t, blobs_t) = self._shell_arrays()
        points, log_l, blobs = shells[shell]
        store('points_{}'.format(shell), points)
        store('log_l_{}'.format(shell), log_l)
        store('blobs_{}'.format(shell), blobs)
        store('points_t', points_t)
        store('shell_t', self._t_shell)
        store('log_l_t', log_l_t)
        store('blobs_t', blobs_t)
        if isinstance(self.bounds[shell], NautilusBound):
            self.bounds[shell].update(fstream['bound_{}'.format(shell)])
        self._write_rng(group)
        fstream.close()
        self._t_checkpoint = time()

    def _resume(self, filepath):
        """Continue from a checkpoint (sampler.py:330-371).  The stored
        points go back into the device arena: first what every shell held
        when the exploration phase ended ('shell_end_exp'), then the rest, so
        that ``discard_exploration`` remains a test on arena positions;
        chronology inside a shell is kept."""
        with open_store(filepath, 'r') as fstream:
            group = fstream['sampler']
            self.rng.bit_generator.state = dict(
                bit_generator='PCG64',
                state=dict(state=int(group.attrs['rng_state']),
                           inc=int(group.attrs['rng_inc'])),
                has_uint32=int(group.attrs['rng_has_uint32']),
                uinteger=int(group.attrs['rng_uinteger']))
            for key in self._ATTRS_STATE:
                if key not in group.attrs:      # written before run()
                    continue
                value = group.attrs[key]
                if np.ndim(value) == 0:
                    value = value.item() if hasattr(value, 'item') else value
                else:
                    value = np.array(value)
                setattr(self, key, value)
            self.explored = bool(self.explored)
            self._discard_exploration = bool(self._discard_exploration)
            if 'p_shell' in group.attrs:
                self._p_shell = {int(i): (int(got), int(raw)) for i, got, raw
                                 in np.array(group.attrs['p_shell']).reshape(
                                     -1, 3)}
            self.shell_n = self.shell_n.astype(int)
            self.shell_n_sample = self.shell_n_sample.astype(int)
            n_shells = len(self.shell_n)

            points = [np.array(group['points_{}'.format(i)], dtype=float)
                      .reshape(-1, self.n_dim) for i in range(n_shells)]
            log_l = [np.array(group['log_l_{}'.format(i)], dtype=float)
                     for i in range(n_shells)]
            blobs = None
            if 'blobs_0' in group:
                blobs = [np.array(group['blobs_{}'.format(i)])
                         for i in range(n_shells)]
                self.blobs_dtype = blobs[0].dtype
            points_t = np.array(group['points_t'], dtype=float).reshape(
                -1, self.n_dim)
            log_l_t = np.array(group['log_l_t'], dtype=float)
            shell_t = np.array(group['shell_t']).astype(int)
            blobs_t = np.array(group['blobs_t']) if 'blobs_t' in group \
                else None

            self.bounds = [UnitCube.read(fstream['bound_0'], rng=self.rng)]
            for i in range(1, n_shells):
                self.bounds.append(NautilusBound.read(
                    fstream['bound_{}'.format(i)], rng=self.rng,
                    mode=self.mlp_mode))

        self._arena = _Arena(self.n_dim, default_device())
        if self.explored:
            end = np.minimum(np.asarray(self.shell_end_exp, dtype=int),
                             [len(p) for p in points])
        else:
            end = np.array([len(p) for p in points], dtype=int)
        parts = [(i, slice(0, end[i])) for i in range(n_shells)] + \
                [(i, slice(end[i], None)) for i in range(n_shells)]
        blob_parts = []
        for n_part, (i, part) in enumerate(parts):
            if n_part == n_shells:
                self._explore_end = self._arena.n if self.explored else 0
            if len(points[i][part]) == 0:
                continue
            self._arena.append(points[i][part], log_l[i][part], i)
            if blobs is not None:
                blob_parts.append(blobs[i][part])
        if not self.explored:
            self._explore_end = 0
        # transfer candidates: stored in the arena, tagged with their donor
        self._t_pos = self._arena.n + np.arange(len(points_t), dtype=np.int64)
        self._t_shell = shell_t
        for s in np.unique(shell_t):
            sel = shell_t == s
            self._arena.append(points_t[sel], log_l_t[sel], -2 - int(s))
            if blobs is not None and blobs_t is not None:
                blob_parts.append(blobs_t[sel])
        if len(shell_t):
            # (appended donor by donor: positions follow that order)
            order = np.argsort(shell_t, kind='stable')
            inverse = np.empty_like(order)
            inverse[order] = np.arange(len(order))
            self._t_pos = self._t_pos[0] + inverse.astype(np.int64)
        if blobs is not None:
            self._blobs_all = (np.concatenate(blob_parts) if blob_parts
                               else np.zeros(0, dtype=self.blobs_dtype))
        self._sums = [None] * n_shells
        self._stack = None
        self._t_checkpoint = time()

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
