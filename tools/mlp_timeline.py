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
buf = np.zeros(512, dtype=np.int64)
for rep in range(3):
    stack.cycle(0, 1 << 20, seed=rep, like_id=like.like_id, like_params=par,
                mode=ops.MLP_TF32)
    torch.cuda.synchronize()
    n = lib.nb200_debug_timeline(buf.ctypes.data_as(ctypes.c_void_p), 512)
tags, t = buf[0:n:2], buf[1:n:2]
names = {1: 'tile start', 2: 'A0 stored+sync', 3: 'L0(0) issued',
         4: 'L0(0) done', 5: 'epi0(0) done', 10: 'loop top',
         11: 'L1 done (wait)', 12: 'L0(n+1) issued', 13: 'epi1 done',
         14: 'L2 issued', 15: 'L0(n+1) done (wait)', 16: 'epi0(n+1) done',
         17: 'L2 done (wait)', 18: 'L1(n+1) issued', 19: 'epi2+dot done'}
prev = t[0]
for tag, ts in list(zip(tags, t))[:70]:
    print('{:5d} +{:6d}  {}'.format(int(ts - t[0]), int(ts - prev),
                                    names.get(int(tag), tag)))
    prev = ts
starts = t[tags == 1]
print('cycles per tile (group 0, CTA 0):', np.diff(starts)[:8])
