"""Stage times of the cycle with L later bounds (grouped exclusion vs loop)."""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from nautilus_b200 import likelihoods, ops  # noqa: E402
from nautilus_b200._pack import flat_to_spec  # noqa: E402

g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests',
                              'golden', 'cfg2_bound_d30.npz')))
spec = flat_to_spec(g)


def nested(f, dthr):
    sp = copy.deepcopy(spec)
    for mx in sp['mixtures']:
        mx['ell']['B'] = mx['ell']['B'] * f
        mx['ell']['B_inv'] = mx['ell']['B_inv'] / f
    for nbs in sp['neural']:
        nbs['ell']['B'] = nbs['ell']['B'] * f
        nbs['ell']['B_inv'] = nbs['ell']['B_inv'] / f
        nbs['score_predict_min'] += dthr
    return sp


L = int(sys.argv[1]) if len(sys.argv) > 1 else 47
mode = {'tf32': ops.MLP_TF32, 'f16': ops.MLP_F16}[
    sys.argv[2] if len(sys.argv) > 2 else 'tf32']
rs = np.random.default_rng(47)
later = [nested(0.9885**(i + 1), 0.01 * rs.normal()) for i in range(L)]
like = likelihoods.Gaussian(30)
par = like.device_params('cuda')
n = 1 << 20
for which in ('grouped', 'loop'):
    if which == 'loop':
        os.environ['NB200_EXCLUDE'] = 'loop'
    else:
        os.environ.pop('NB200_EXCLUDE', None)
    stack = ops.DeviceStack([spec] + later)
    out = stack.cycle(0, n, later=(1, L), like_id=like.like_id,
                      like_params=par, mode=mode)
    torch.cuda.synchronize()
    ops.profile_enable(True)
    for s in range(3):
        stack.cycle(0, n, later=(1, L), offset=s * n, like_id=like.like_id,
                    like_params=par, mode=mode, out=out)
    prof = ops.profile_collect()
    ops.profile_enable(False)
    print(which, {k: round(v[0] / 3, 3) for k, v in prof.items() if v[1]},
          out['counters'].cpu().numpy())
