"""Cycle-level timeline of k_mlp_tf32 (leader thread of tile group 0, CTA 0).
Needs a library built with -DNB200_TIMELINE:
    NB200_EXTRA_FLAGS=-DNB200_TIMELINE python tools/mlp_timeline.py"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
os.environ.setdefault('NB200_EXTRA_FLAGS', '-DNB200_TIMELINE')

from nautilus_b200 import _lib  # noqa: E402
_lib.build(force=True)
import torch  # noqa: E402
from nautilus_b200 import likelihoods, ops  # noqa: E402
from nautilus_b200._pack import flat_to_spec  # noqa: E402

with np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden',
                          'cfg2_bound_d30.npz')) as f:
    spec = flat_to_spec({k: f[k] for k in f.files})
stack = ops.DeviceStack([spec])
like = likelihoods.Gaussian(30)
par = like.device_params('cuda')
lib = _lib.lib()
lib.nb200_debug_timeline.restype = ctypes.c_int
buf = np.zeros(2048, dtype=np.int64)
for rep in range(3):
    stack.cycle(0, 1 << 20, seed=rep, like_id=like.like_id, like_params=par,
                mode=ops.MLP_TF32)
    torch.cuda.synchronize()
    n = lib.nb200_debug_timeline(buf.ctypes.data_as(ctypes.c_void_p), 2048)
names = {1: 'tile staged', 10: 'loop top', 11: 'L1 done (wait)',
         13: 'epi1 done', 15: 'L0(i+1) done (wait)', 16: 'epi0(i+1) done',
         17: 'L2 done (wait)', 19: 'epi2+dot done',
         20: 'issuer: epi0 seen', 21: 'issuer: L1(i), L0(i+1) issued',
         22: 'issuer: epi1 + last epi seen', 23: 'issuer: L2 issued'}
for slot, who in ((0, 'epilogue thread 0'), (1, 'issuer of group 0')):
    row = buf[slot * 1024:(slot + 1) * 1024]
    tags, t = row[0::2], row[1::2]
    k = int(np.count_nonzero(t))
    tags, t = tags[:k], t[:k]
    print('==', who, '({} stamps)'.format(k))
    prev = t[0] if k else 0
    for tag, ts in list(zip(tags, t))[40:120]:
        print('{:7d} +{:6d}  {}'.format(int(ts - t[0]), int(ts - prev),
                                        names.get(int(tag), tag)))
        prev = ts
    if slot == 0 and k:
        starts = t[tags == 1]
        print('cycles per tile (group 0, CTA 0):', np.diff(starts)[:10])
