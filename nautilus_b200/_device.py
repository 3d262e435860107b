"""Plumbing shared by the bound classes: array conversion and Philox streams."""

import numpy as np
import torch


_HAVE_CUDA = None


def default_device():
    global _HAVE_CUDA
    if _HAVE_CUDA is None:          # asked thousands of times per run
        _HAVE_CUDA = bool(torch.cuda.is_available())
    if not _HAVE_CUDA:
        raise RuntimeError(
            'nautilus_b200 needs a CUDA device (B200); there is no CPU path.')
    return torch.device('cuda', torch.cuda.current_device())


def to_device(points, n_dim):
    """Return (contiguous CUDA float64 [n, d] tensor, restore) where restore
    maps a per-point device result back to the caller's array type/shape."""
    if isinstance(points, torch.Tensor):
        t = points
        kind = 'torch'
    else:
        t = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64))
        kind = 'numpy'
    single = t.dim() == 1
    if single:
        t = t.unsqueeze(0)
    if t.dim() != 2 or t.shape[1] != n_dim:
        raise ValueError('points must have shape (n, {})'.format(n_dim))
    t = t.to(device=default_device(), dtype=torch.float64).contiguous()

    def restore(res):
        if single:
            res = res[0]
        if kind == 'numpy':
            res = res.cpu().numpy()
            if single and res.ndim == 0:
                res = res[()]
        return res

    return t, restore


class PhiloxStream:
    """The random stream of one bound: (seed, stream id, running offset).

    Stands in for the ``numpy.random.Generator`` the reference hands to every
    bound (nautilus/sampler.py:1002,1031): the seed is drawn from that
    generator, so runs are reproducible from the sampler seed; ``take(n)``
    reserves n proposal indices."""

    def __init__(self, rng=None):
        self.reseed(rng)

    def reseed(self, rng=None):
        # seed and stream id are both drawn from the caller's generator: two
        # bounds built from identically seeded generators sample identically,
        # two bounds built one after the other from one generator do not
        if rng is None:
            rng = np.random.default_rng()
        if isinstance(rng, (int, np.integer)):
            rng = np.random.default_rng(int(rng))
        self.seed = int(rng.integers(0, 2**63 - 1))
        self.stream_id = int(rng.integers(0, 2**32 - 1))
        self.offset = 0

    def take(self, n):
        start = self.offset
        self.offset += int(n)
        return start

    # -- checkpoints: the stream is three integers -------------------------
    def write(self, group):
        """Stands in for the PCG64 state the reference stores
        (nautilus/sampler.py:1325-1329)."""
        group.attrs['philox_seed'] = np.int64(self.seed)
        group.attrs['philox_stream_id'] = np.int64(self.stream_id)
        group.attrs['philox_offset'] = np.int64(self.offset)

    @classmethod
    def read(cls, group, rng=None):
        """The stored stream, or a fresh one drawn from ``rng`` when the
        group was written by the reference (no Philox attributes)."""
        if 'philox_seed' not in group.attrs:
            return cls(rng)
        stream = cls.__new__(cls)
        stream.seed = int(group.attrs['philox_seed'])
        stream.stream_id = int(group.attrs['philox_stream_id'])
        stream.offset = int(group.attrs['philox_offset'])
        return stream
