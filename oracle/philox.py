"""Philox4x32-10 counter RNG in NumPy (oracle side of the kernels' RNG).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference draws from one serial PCG64 stream (``nautilus/sampler.py:305``)
which a data-parallel sampler cannot reproduce; the CUDA kernels use the
counter-based Philox4x32-10 of Salmon et al. (SC'11, "Parallel random numbers:
as easy as 1, 2, 3") keyed by the sampler seed with the *global proposal index*
as counter, so results do not depend on how the batch is sharded.  This file
restates the published algorithm so the integer-exact parts of the stream (the
ellipsoid choice, the overlap-acceptance uniform, the radial uniform, cube
uniforms) can be replayed bit-for-bit on the CPU.

Counter layout used by ``nautilus_b200/csrc`` (``nb200_rng.cuh``):
    ctr = (idx_lo, idx_hi, block, stream)    key = (seed_lo, seed_hi)
``idx`` = global proposal index (call offset included), ``block`` = running
block number inside one proposal (each block yields 4 x uint32), ``stream`` =
tag separating bounds / purposes.
"""

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Return the 4 output words for arrays of counters and a scalar key."""
    c0 = np.asarray(c0, dtype=np.uint64) & MASK
    c1 = np.asarray(c1, dtype=np.uint64) & MASK
    c2 = np.asarray(c2, dtype=np.uint64) & MASK
    c3 = np.asarray(c3, dtype=np.uint64) & MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c0
            p1 = M1 * c2
            hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
            n0 = hi1 ^ c1 ^ np.uint64(k0)
            n1 = lo1
            n2 = hi0 ^ c3 ^ np.uint64(k1)
            n3 = lo0
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def philox_block(idx, block, stream, seed):
    """4 words for proposal index array ``idx`` (uint64) and a block number."""
    idx = np.asarray(idx, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return philox4x32_10(idx & MASK, idx >> np.uint64(32),
                         np.uint64(block), np.uint64(stream),
                         seed & 0xFFFFFFFF, seed >> 32)


def u01_32(w):
    """(w + 0.5) * 2^-32 as float64: uniform in (0, 1), exact in fp64."""
    return (w.astype(np.float64) + 0.5) * 2.0**-32


def u01_53(wa, wb):
    """53-bit uniform in [0, 1) from two words (27 + 26 bits), exact."""
    a = (wa >> np.uint32(5)).astype(np.float64)
    b = (wb >> np.uint32(6)).astype(np.float64)
    return (a * 67108864.0 + b) * 2.0**-53
