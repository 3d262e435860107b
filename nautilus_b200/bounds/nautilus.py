"""The composite nautilus bound: union of cube-ellipsoid mixtures AND any of
the neural-network bounds.

Host-side mirror of ``nautilus/bounds/nautilus.py``; ``PhaseShift``
(periodic parameters) is outside the hot-path scope
and rejected loudly.
"""

import numpy as np
import torch

from .. import ops
from ..neural import NeuralNetworkEmulator
from .._device import PhiloxStream, default_device, to_device
from .basic import Ellipsoid, UnitCubeEllipsoidMixture, _DeviceBound
from .neural import NeuralBound
from .union import Union


_SIDE_STREAMS = {}


def _side_stream(device, which=0):
    """Side streams of a device for the emulator fits, kept for the life of
    the process: a new stream per bound starts with an empty allocator pool,
    and every tensor made under it is a cudaMalloc."""
    key = (device.type, device.index, which)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class NautilusBound(_DeviceBound):
    """(nautilus/bounds/nautilus.py:13-397)."""

    raw_batch = 1 << 16

    @classmethod
    def compute(cls, points, log_l, log_l_min, log_v_target,
                enlarge_per_dim=1.1, n_points_min=None, split_threshold=100,
                periodic=None, n_networks=4, neural_network_kwargs={},
                pool=None, rng=None, mode=None):
        """``mode`` selects the emulator arithmetic of this bound
        (``ops.MLP_F64`` parity mode / ``ops.MLP_TF32`` tensor cores); the
        same arithmetic is used for sampling and for ``contains``."""
        if periodic is not None:
            raise NotImplementedError(
                'periodic parameters (PhaseShift) are outside the scope of '
                'nautilus_b200.')
        # `points` / `log_l` may be CUDA tensors (the Sampler keeps every
        # evaluated point on the device): only the live points and the
        # training subsets are brought to the host, never the whole set
        on_device = isinstance(points, torch.Tensor)
        if not on_device:
            points = np.asarray(points, dtype=float)
            log_l = np.asarray(log_l, dtype=float)
        bound = cls()
        bound.n_dim = points.shape[1]
        bound.shift = None
        bound.mode = NeuralNetworkEmulator.mode if mode is None else mode
        bound.rng = np.random.default_rng() if rng is None else rng
        live = points[log_l >= log_l_min]
        if on_device:
            live = live.cpu().numpy()

        # one neural bound per non-overlapping live-point ellipsoid
        clusters = Union.compute(
            live, enlarge_per_dim=enlarge_per_dim, n_points_min=n_points_min,
            bound_class=Ellipsoid, rng=rng)
        # The emulator fits (32 SMs, ~0.1 s each) run on a side stream while
        # the host goes on building bounds on the main stream.  The fit of the
        # UNSPLIT live-point ellipsoid is started speculatively, before the
        # split attempt (mixture EM + two enclosing ellipsoids, ~60 ms of
        # host-driven work): most bounds of a unimodal problem keep the single
        # ellipsoid, and the fit then overlaps the split attempt AND the outer
        # bound instead of the outer bound alone.  If the split succeeds the
        # speculative fit is dropped (it finishes in the background on its own
        # stream) and one fit per cluster is enqueued as before.
        main = torch.cuda.current_stream()
        side = _side_stream(main.device, 0)
        spec_stream = _side_stream(main.device, 1)
        nb_kwargs = dict(enlarge_per_dim=enlarge_per_dim,
                         n_networks=n_networks,
                         neural_network_kwargs=neural_network_kwargs,
                         pool=pool, rng=rng, mode=bound.mode, defer=True)
        first = clusters.bounds[0]
        spec_stream.wait_stream(main)
        with torch.cuda.stream(spec_stream):
            member = first.contains(points)
            speculative = NeuralBound.compute(points[member], log_l[member],
                                              log_l_min, **nb_kwargs)
        # (device copies made under that stream stay with that stream)
        first._invalidate()
        first._dev_params = None
        while clusters.split(allow_overlap=False):
            pass
        if len(clusters.bounds) == 1 and clusters.bounds[0] is first:
            bound.neural_bounds = [speculative]
            side = spec_stream
        else:
            del speculative
            bound.neural_bounds = []
            side.wait_stream(main)
            with torch.cuda.stream(side):
                for ell in clusters.bounds:
                    member = ell.contains(points)
                    bound.neural_bounds.append(NeuralBound.compute(
                        points[member], log_l[member], log_l_min,
                        **nb_kwargs))

        # outer sampling bound, refined until close enough to the target volume
        bound.outer_bound = Union.compute(
            live, enlarge_per_dim=enlarge_per_dim, n_points_min=n_points_min,
            bound_class=UnitCubeEllipsoidMixture, rng=rng)
        slack = np.log(split_threshold * enlarge_per_dim**bound.n_dim)
        while bound.outer_bound.log_v - log_v_target > slack:
            if not bound.outer_bound.split():
                break
        while bound.outer_bound.log_v - log_v_target > slack:
            if not bound.outer_bound.trim():
                break

        with torch.cuda.stream(side):
            for nb in bound.neural_bounds:
                nb.finish()
        main.wait_stream(side)
        # device copies made under the side stream (ellipsoid parameters,
        # serialised stacks) belong to that stream's allocator pool: drop them
        # so that whatever the bound uses from now on is created on the
        # caller's stream (no cross-stream reuse of freed blocks)
        for nb in bound.neural_bounds:
            nb._invalidate()
            nb.outer_bound._invalidate()
            nb.outer_bound._dev_params = None
            if nb.emulator is not None:
                nb.emulator._stack = None
        for ell in clusters.bounds:
            ell._invalidate()
            ell._dev_params = None

        bound.stream = PhiloxStream(rng)
        bound._clear()
        return bound

    def _clear(self):
        self.points = np.zeros((0, self.n_dim))
        self._buffer = None
        self.n_sample = 0
        self.n_reject = 0
        self._replicas = None
        self._invalidate()

    def spec(self):
        spec = self.outer_bound.spec()
        spec['neural'] = [nb.nb_spec() for nb in self.neural_bounds]
        return spec

    def contains(self, points, mode=None):
        """outer union AND any neural bound (nautilus.py:146-169)."""
        t, restore = to_device(points, self.n_dim)
        mode = self.mode if mode is None else mode
        return restore(self._device_stack().contains(0, t, mode=mode))

    # -- sampling ------------------------------------------------------------
    def draw_raw(self, n_raw, mode=None, pool=None):
        """One raw batch through union proposal and neural filter.  Updates
        all four integer counters exactly as the reference's nested loops do
        in aggregate (union.py:322-323, nautilus.py:221-222) and returns the
        accepted points (CUDA).

        With a ``GpuPool`` the batch is sharded over its devices the way the
        reference shards it over worker processes (nautilus.py:223-237):
        device i draws the global proposal indices of its slice (Philox is
        keyed by the global index, so the result does not depend on the pool
        size), the parent sums the four counters and concatenates the points
        on the first device."""
        mode = self.mode if mode is None else mode
        n_raw = int(n_raw)
        offset = self.stream.take(n_raw)
        home = self._device_stack().device     # where the FIFO lives
        if pool is None or getattr(pool, 'size', 1) <= 1 or \
                not hasattr(pool, 'devices'):
            devices, slices = [home], [(0, n_raw)]
        else:
            devices, slices = pool.devices, pool.slices(n_raw)
        launched = []
        for dev, (lo, hi) in zip(devices, slices):
            if hi <= lo:
                continue
            with torch.cuda.device(dev):
                stack = self._device_stack_on(dev)
                out = stack.cycle(0, hi - lo, seed=self.stream.seed,
                                  offset=offset + lo,
                                  stream_id=self.stream.stream_id, mode=mode)
                keep, _, _ = stack.compact(out['points'], None, out['code'])
                launched.append((keep, out['counters']))
        parts, cnt = [], np.zeros(ops.N_CNT, dtype=np.int64)
        for keep, counters in launched:          # one sync per device
            c = counters.cpu().numpy()
            cnt += c
            parts.append(keep[:int(c[ops.CNT_IN_SHELL])].to(home))
        n_union_reject = int(cnt[ops.CNT_CUBE_REJECT] +
                             cnt[ops.CNT_OVERLAP_REJECT])
        self.outer_bound.n_sample += n_raw
        self.outer_bound.n_reject += n_union_reject
        self.n_sample += n_raw - n_union_reject
        self.n_reject += int(cnt[ops.CNT_NN_REJECT])
        return torch.cat(parts) if len(parts) > 1 else parts[0]

    def _device_stack_on(self, dev):
        """The serialised bound on device ``dev`` (replicated on demand, like
        the reference pickles the bound to every worker)."""
        dev = torch.device(dev)
        if dev == self._device_stack().device:
            return self._device_stack()
        if getattr(self, '_replicas', None) is None:
            self._replicas = {}
        if dev not in self._replicas:
            self._replicas[dev] = ops.DeviceStack([self.spec()], device=dev)
        return self._replicas[dev]

    def _refill(self, n_points, pool=None):
        have = 0 if self._buffer is None else self._buffer.shape[0]
        chunks = [] if self._buffer is None else [self._buffer]
        while have < n_points:
            if self.outer_bound.n_sample > 0 and self.n_sample > 0:
                acc = ((1 - self.outer_bound.n_reject /
                        self.outer_bound.n_sample) *
                       (1 - self.n_reject / self.n_sample))
            else:
                acc = 0.25
            n_raw = int(min(max(self.raw_batch,
                                1.2 * (n_points - have) / max(acc, 1e-4)),
                            1 << 22))
            keep = self.draw_raw(n_raw, pool=pool)
            chunks.append(keep)
            have += keep.shape[0]
        self._buffer = torch.cat(chunks) if len(chunks) > 1 else chunks[0]

    def sample(self, n_points=100, return_points=True, pool=None,
               as_numpy=True):
        """Pop n points from the FIFO of accepted points, refilling it with
        raw GPU batches (nautilus.py:193-244).  ``pool`` may be a ``GpuPool``:
        the raw batch is then sharded over its devices."""
        self._refill(n_points, pool=pool)
        if not return_points:
            return None
        out = self._buffer[:n_points]
        self._buffer = self._buffer[n_points:]
        return out.cpu().numpy() if as_numpy else out.contiguous()

    @property
    def log_v(self):
        """outer.log_v + log(1 - n_reject / n_sample) (nautilus.py:246-261)."""
        if self.n_sample == 0:
            self.sample(return_points=False)
        return self.outer_bound.log_v + np.log(
            1.0 - self.n_reject / self.n_sample)

    @property
    def n_ell(self):
        """Number of ellipsoids (nautilus.py:263-274)."""
        return int(np.sum([np.any(~b.dim_cube)
                           for b in self.outer_bound.bounds]))

    @property
    def n_net(self):
        """Number of networks (nautilus.py:276-290)."""
        emu = self.neural_bounds[0].emulator
        if emu is None:
            return 0
        return len(self.neural_bounds) * len(emu.neural_networks)

    # -- checkpoints --------------------------------------------------------
    def _fifo_host(self):
        if self._buffer is None:
            return np.zeros((0, self.n_dim))
        return self._buffer.cpu().numpy()

    def write(self, group):
        """Layout of nautilus/bounds/nautilus.py:292-317 plus the Philox
        stream of the bound."""
        group.attrs['type'] = 'NautilusBound'
        group.attrs['n_dim'] = self.n_dim
        group.attrs['n_neural_bounds'] = len(self.neural_bounds)
        for i, neural_bound in enumerate(self.neural_bounds):
            neural_bound.write(group.create_group('neural_bound_{}'.format(i)))
        self.outer_bound.write(group.create_group('outer_bound'))
        group.create_dataset('points', data=self._fifo_host(),
                             maxshape=(None, self.n_dim))
        group.attrs['n_sample'] = self.n_sample
        group.attrs['n_reject'] = self.n_reject
        self.stream.write(group)

    def update(self, group):
        """(nautilus.py:319-333)."""
        group.attrs['n_sample'] = self.n_sample
        group.attrs['n_reject'] = self.n_reject
        self.outer_bound.update(group['outer_bound'])
        fifo = self._fifo_host()
        group['points'].resize(fifo.shape)
        group['points'][...] = fifo
        self.stream.write(group)

    @classmethod
    def read(cls, group, rng=None, mode=None):
        """(nautilus.py:335-380).  ``mode``: emulator arithmetic of the bound
        (as in ``compute``).  A group written by the reference has no Philox
        attributes: the streams are then drawn from ``rng``."""
        if 'shift' in group:
            raise NotImplementedError(
                'periodic parameters (PhaseShift) are outside the scope of '
                'nautilus_b200.')
        bound = cls()
        bound.rng = np.random.default_rng() if rng is None else rng
        bound.mode = NeuralNetworkEmulator.mode if mode is None else mode
        bound.n_dim = int(group.attrs['n_dim'])
        bound.shift = None
        bound.neural_bounds = []
        while 'neural_bound_{}'.format(len(bound.neural_bounds)) in group:
            bound.neural_bounds.append(NeuralBound.read(
                group['neural_bound_{}'.format(len(bound.neural_bounds))],
                rng=bound.rng, mode=bound.mode))
        bound.outer_bound = Union.read(group['outer_bound'], rng=bound.rng)
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
        """Forget sampling progress; optionally reseed (nautilus.py:382-397)."""
        self._buffer = None
        self.points = np.zeros((0, self.n_dim))
        self.n_sample = 0
        self.n_reject = 0
        self.outer_bound.reset(rng)
        if rng is not None:
            self.rng = rng
            self.stream.reseed(rng)
