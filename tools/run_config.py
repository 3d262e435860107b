"""End-to-end runs of the BASELINE.json configurations through the drop-in
Sampler; prints one JSON line per run (wall time, n_like, bounds, log Z vs the
analytic truth, N_eff).  Usage: python tools/run_config.py --config 2"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))

from nautilus_b200 import Sampler, likelihoods  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', type=int, default=2)
    ap.add_argument('--n-eff', type=float, default=10000)
    ap.add_argument('--n-batch', type=int, default=None)
    ap.add_argument('--arith', default='tf32')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--keep-exploration', action='store_true')
    ap.add_argument('--timeout', type=float, default=1500)
    ap.add_argument('--profile', action='store_true')
    args = ap.parse_args()
    nn_kwargs = {}
    if args.config == 1:
        like, n_live = likelihoods.Gaussian(3, mu=[0.4, 0.5, 0.6],
                                            sigma=0.1), 1000
    elif args.config == 2:
        like, n_live = likelihoods.Gaussian(30, sigma=0.1), 2000
    elif args.config == 3:
        like, n_live = likelihoods.Rosenbrock(50), 4000
        nn_kwargs = dict(hidden_layer_sizes=(128, 128, 128, 128))
    elif args.config == 4:
        mus = np.full((4, 30), 0.5)
        mus[:, 0] = [0.25, 0.25, 0.75, 0.75]
        mus[:, 1] = [0.25, 0.75, 0.25, 0.75]
        like, n_live = likelihoods.GaussianMixture(mus, sigma=0.03), 2000
    else:
        raise SystemExit('config not wired')
    sampler = Sampler(lambda x: x, like, n_dim=like.n_dim, n_live=n_live,
                      seed=args.seed, n_batch=args.n_batch,
                      emulator_arith=args.arith,
                      neural_network_kwargs=nn_kwargs)
    t0 = time.time()
    if args.profile:
        import cProfile
        import pstats
        prof = cProfile.Profile()
        prof.enable()
    ok = sampler.run(n_eff=args.n_eff, timeout=args.timeout,
                     discard_exploration=not args.keep_exploration)
    wall = time.time() - t0
    if args.profile:
        prof.disable()
        st = pstats.Stats(prof).sort_stats('cumulative')
        st.print_stats(28)
        st.print_callers('is_available|torch.empty|linalg_inv')
    raw = sum(b.outer_bound.n_sample for b in sampler.bounds[1:])
    emus = [nb.emulator for b in sampler.bounds[1:] for nb in b.neural_bounds
            if nb.emulator is not None]
    fit = dict(n_fits=len(emus),
               rows_mean=float(np.mean([e.n_train_ for e in emus])),
               rows_max=int(np.max([e.n_train_ for e in emus])),
               epochs_mean=float(np.mean([n.n_iter_ for e in emus
                                          for n in e.neural_networks])),
               epochs_max=int(np.max([n.n_iter_ for e in emus
                                      for n in e.neural_networks]))) \
        if emus else None
    print(json.dumps({
        'config': args.config, 'success': bool(ok), 'wall_s': wall,
        'n_like': int(sampler.n_like), 'n_bounds': len(sampler.bounds),
        'log_z': float(sampler.log_z), 'log_z_true': like.log_z_true,
        'delta_log_z': (None if like.log_z_true is None else
                        abs(float(sampler.log_z) - like.log_z_true)),
        'f_live': None if sampler.explored else float(sampler.f_live),
        'n_eff': float(sampler.n_eff), 'raw_proposals': int(raw),
        'emulator_fits': fit, 'emulator_arith': args.arith, 'n_batch': sampler.n_batch,
        'discard_exploration': not args.keep_exploration}), flush=True)


if __name__ == '__main__':
    main()
