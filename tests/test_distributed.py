"""The N > 1 path on CPU: two gloo ranks exchange per-rank counters and
log-sum-exp partials with one collective and must reproduce the single-rank
answer (the analogue of the reference's pool test, tests/test_pool.py, and of
tests/test_bounds.py:412-441: results do not depend on the number of workers).
"""

import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from nautilus_b200.pool import (GpuPool, exchange_packed_async,  # noqa: E402
                                exchange_stats, merge_lse, merge_packed,
                                packed_stats)
from oracle import nautilus_oracle as orc  # noqa: E402
from oracle import philox  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, n, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    # every rank owns a contiguous slice of the global proposal index space;
    # the "likelihood" is a deterministic function of the Philox stream so
    # the union over ranks is exactly the single-rank batch
    lo, hi = n * rank // world, n * (rank + 1) // world
    idx = np.arange(lo, hi, dtype=np.uint64)
    w = philox.philox_block(idx, 0, 7, 123)
    log_l = -50 * philox.u01_32(w[0])
    code = (w[1] % 5).astype(np.uint8)
    sel = code == 4
    cnt = np.zeros(8, dtype=np.int64)
    cnt[0] = len(idx)
    for c in range(4):
        cnt[1 + c] = np.sum(code == c)
    cnt[5] = np.sum(sel)
    cnt[6] = np.sum(log_l[sel] >= -10)
    m, s1, s2 = orc.lse_triple(log_l[sel])
    total, lse = exchange_stats(torch.from_numpy(cnt),
                                torch.tensor([m, s1, s2, 0.0],
                                             dtype=torch.float64))
    # the packed form (what bench.py uses: kernels write into the send
    # buffer) must give the same answer, with exact int64 counters
    words, lse_v, cnt_v = packed_stats('cpu')
    cnt_v.copy_(torch.from_numpy(cnt))
    lse_v.copy_(torch.tensor([m, s1, s2, 0.0], dtype=torch.float64))
    total2, lse2 = merge_packed(exchange_packed_async(words))
    assert total2.tolist() == total.tolist() and lse2 == lse
    # the asynchronous, double-buffered form of bench.py: the handle of step
    # s is waited for only when step s + 2 reuses the send buffer
    bufs = [packed_stats('cpu') for _ in range(2)]
    gath = [torch.zeros((world, 12), dtype=torch.int64) for _ in range(2)]
    works = [None, None]
    for s in range(5):
        b = s & 1
        if works[b] is not None:
            works[b].wait()
            t3, l3 = merge_packed(gath[b])
            assert t3.tolist() == (total * (s - 1)).tolist()
        bufs[b][2].copy_(torch.from_numpy(cnt * (s + 1)))
        bufs[b][1].copy_(torch.tensor([m, s1, s2, 0.0], dtype=torch.float64))
        works[b] = exchange_packed_async(bufs[b][0], gathered=gath[b],
                                         async_op=True)
    for b in range(2):
        works[b].wait()
    t4, l4 = merge_packed(gath[0])         # step 4 used buffer 0
    assert t4.tolist() == (total * 5).tolist() and l4 == lse
    if rank == 0:
        out.put((total.tolist(), lse))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_exchange_matches_single_rank(world):
    n = 100003
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, n, out))
             for r in range(world)]
    for p in procs:
        p.start()
    total, lse = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    idx = np.arange(n, dtype=np.uint64)
    w = philox.philox_block(idx, 0, 7, 123)
    log_l = -50 * philox.u01_32(w[0])
    code = (w[1] % 5).astype(np.uint8)
    sel = code == 4
    assert total[0] == n and total[5] == np.sum(sel)
    assert total[1:5] == [int(np.sum(code == c)) for c in range(4)]
    assert total[6] == np.sum(log_l[sel] >= -10)
    m, s1, s2 = orc.lse_triple(log_l[sel])
    assert lse[0] == m
    assert abs(lse[1] / s1 - 1) < 1e-12 and abs(lse[2] / s2 - 1) < 1e-12


def test_merge_lse_edge_cases():
    assert merge_lse([(-np.inf, 0, 0), (-np.inf, 0, 0)])[0] == -np.inf
    m, s1, s2 = merge_lse([(-np.inf, 0, 0), (2.0, 3.0, 1.5)])
    assert (m, s1, s2) == (2.0, 3.0, 1.5)
    m, s1, s2 = merge_lse([(0.0, 1.0, 1.0), (np.log(2.0), 1.0, 1.0)])
    assert np.isclose(m + np.log(s1), np.log(3.0))
    assert np.isclose(s1 * s1 / s2, 9 / 5)


def test_gpu_pool_slices():
    class FakePool(GpuPool):
        def __init__(self, k):
            self.devices = list(range(k))
    p = FakePool(3)
    assert p.size == 3
    assert p.slices(10) == [(0, 3), (3, 6), (6, 10)]
    assert p.slices(2) == [(0, 0), (0, 1), (1, 2)]
