"""Pools: how the proposal batch and the likelihood calls are spread out.

The reference's only "distributed backend" is ``NautilusPool``
(nautilus/pool.py:36-107): a ``.map/.size`` facade over multiprocessing, dask
or MPI executors, used for likelihood calls (sampler.py:863-873), for proposal
sampling (bounds/nautilus.py:223-237) and for network training
(nautilus/neural.py:93-96).  Here the proposal batch is sharded over GPUs
instead: ``GpuPool`` names the devices (one CUDA stream each), every device
draws its slice of the global proposal index range -- Philox counters are
keyed by that index, so results do not depend on ``size`` -- and the per-rank
counters / log-sum-exp partials are merged by ONE collective.  Inside one
process the merge is a host-side sum over devices; across processes
(``torchrun``, one rank per GPU) it is ``torch.distributed.all_gather`` over
NCCL (bench.py, tests/test_distributed.py).

``NautilusPool`` is kept for host-side (Python callable) likelihoods with the
reference's semantics.
"""

from multiprocessing import Pool

_LIKELIHOOD = None


def initialize_worker(likelihood):
    """Cache the likelihood in the worker (nautilus/pool.py:6-16)."""
    global _LIKELIHOOD
    _LIKELIHOOD = likelihood


def likelihood_worker(*args):
    """Evaluate the cached likelihood (nautilus/pool.py:19-33)."""
    return _LIKELIHOOD(*args)


class NautilusPool:
    """``.map`` / ``.size`` facade over executor-like pools
    (nautilus/pool.py:36-107)."""

    def __init__(self, pool, likelihood=None):
        if isinstance(pool, int):
            pool = Pool(pool, initializer=initialize_worker,
                        initargs=(likelihood, ))
        self.pool = pool

    def _is_dask(self):
        return 'distributed.client.Client' in str(type(self.pool))

    def map(self, func, iterable):
        if self._is_dask():
            return list(self.pool.gather(self.pool.map(func, iterable)))
        return list(self.pool.map(func, iterable))

    @property
    def size(self):
        if self._is_dask():
            return len(self.pool.nthreads())
        for attr in ('_processes', '_max_workers', 'size', 'nt'):
            if hasattr(self.pool, attr):
                return getattr(self.pool, attr)
        raise ValueError('Cannot determine size of pool.')


class GpuPool:
    """A set of CUDA devices that share one proposal batch."""

    def __init__(self, devices=None):
        import torch
        if devices is None:
            devices = 1
        if isinstance(devices, int):
            if devices < 1 or devices > torch.cuda.device_count():
                raise ValueError('{} GPUs requested, {} visible.'.format(
                    devices, torch.cuda.device_count()))
            devices = list(range(devices))
        self.devices = [torch.device('cuda', int(i)) for i in devices]

    @property
    def size(self):
        return len(self.devices)

    def slices(self, n):
        """Contiguous [lo, hi) slices of n proposals, one per device."""
        edges = [n * i // self.size for i in range(self.size + 1)]
        return list(zip(edges[:-1], edges[1:]))

    def map(self, func, iterable):
        return [func(item) for item in iterable]


# --------------------------------------------------------------------------
# the one exchange step of the sharded cycle
# --------------------------------------------------------------------------

def merge_lse(parts):
    """Combine per-rank (max, sum e^{l-m}, sum e^{2(l-m)}) triples, in rank
    order so the result is deterministic."""
    import numpy as np
    m = max(float(p[0]) for p in parts)
    if not np.isfinite(m):
        return m, 0.0, 0.0
    s1 = sum(float(p[1]) * np.exp(float(p[0]) - m) for p in parts
             if np.isfinite(p[0]))
    s2 = sum(float(p[2]) * np.exp(2 * (float(p[0]) - m)) for p in parts
             if np.isfinite(p[0]))
    return m, s1, s2


def exchange_stats_async(counters, lse, gathered=None, packed=None,
                         group=None):
    """Enqueue the one collective of a cycle: all-gather every rank's int64
    counters and fp64 LSE partials (NCCL on GPUs, gloo in the CPU tests).

    counters i64[n_cnt], lse f64[>=3] are tensors on this rank's device.
    Counters (< 2^53) travel as float64 in the same buffer as the partials so
    that a single collective suffices.  Returns the gathered [world, n_cnt+4]
    tensor (still on the device, nothing is synchronised); hand it to
    :func:`merge_gathered` when the numbers are needed on the host."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_cnt = counters.numel()
    if packed is None:
        packed = torch.empty(n_cnt + 4, dtype=torch.float64,
                             device=counters.device)
    packed[:n_cnt] = counters.double()
    packed[n_cnt:n_cnt + 3] = lse[:3]
    packed[n_cnt + 3] = 0
    if gathered is None:
        gathered = torch.empty((world, n_cnt + 4), dtype=torch.float64,
                               device=counters.device)
    dist.all_gather_into_tensor(gathered.view(-1), packed, group=group)
    return gathered


def packed_stats(device, n_cnt=8, n_lse=4):
    """One 8-byte-word buffer that holds a cycle's LSE partials (float64) and
    counters (int64) side by side, so that the kernels write straight into
    the send buffer of the collective: returns (words i64[n_lse + n_cnt],
    lse f64 view, counters i64 view)."""
    import torch
    words = torch.zeros(n_lse + n_cnt, dtype=torch.int64, device=device)
    return words, words[:n_lse].view(torch.float64), words[n_lse:]


def exchange_packed_async(words, gathered=None, group=None, async_op=False):
    """All-gather the raw words of :func:`packed_stats` (no packing kernels,
    counters stay exact int64).  Returns the gathered [world, n] int64 tensor
    on the device; nothing is synchronised.

    ``async_op=True`` returns the work handle instead: the collective runs
    on the backend's own stream behind whatever the current stream has
    enqueued so far, and the current stream does NOT wait for it -- call
    ``handle.wait()`` before reusing ``words`` or reading ``gathered`` (with
    two send buffers that is two steps later, i.e. never a stall)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if gathered is None:
        gathered = torch.empty((world, words.numel()), dtype=torch.int64,
                               device=words.device)
    work = dist.all_gather_into_tensor(gathered.view(-1), words, group=group,
                                       async_op=async_op)
    return work if async_op else gathered


def merge_packed(gathered, n_lse=4):
    """(counters int64 ndarray, (m, s1, s2)) from :func:`exchange_packed_async`;
    every rank gets the same answer (merge in rank order)."""
    import numpy as np
    g = np.ascontiguousarray(gathered.cpu().numpy())
    lse = g[:, :n_lse].copy().view(np.float64)
    total = g[:, n_lse:].sum(axis=0)
    return total, merge_lse([tuple(r[:3]) for r in lse])


def merge_gathered(gathered, n_cnt=8):
    """(counters int64 ndarray, (m, s1, s2)) from the gathered per-rank rows;
    every rank gets the same answer (merge in rank order)."""
    g = gathered.cpu().numpy()
    total = g[:, :n_cnt].sum(axis=0).round().astype('int64')
    return total, merge_lse([tuple(r[n_cnt:n_cnt + 3]) for r in g])


def exchange_stats(counters, lse, gathered=None, packed=None, group=None):
    """Blocking form: exchange and merge."""
    return merge_gathered(exchange_stats_async(
        counters, lse, gathered=gathered, packed=packed, group=group),
        n_cnt=counters.numel())
