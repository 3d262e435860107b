"""Time the emulator trainer (tensor-core vs SIMT kernel) on a config-2-like
training set: python tools/bench_fit.py [rows] [d]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from nautilus_b200.neural import NeuralNetworkEmulator  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rng = np.random.default_rng(0)
x = rng.normal(size=(rows, d))
r = np.linalg.norm(x, axis=1)
y = np.argsort(np.argsort(-r)) / rows
for kind in ('tc', 'ffma', 'tc', 'ffma'):
    os.environ['NB200_FIT'] = kind
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    emu = NeuralNetworkEmulator.train(x, y, n_networks=4, seed=3)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ep = [n.n_iter_ for n in emu.neural_networks]
    rmse = np.sqrt(np.mean((emu.predict(x) - y)**2))
    steps = sum(ep) / 4 * np.ceil(rows / 200)
    print('{:5s} {:.3f} s  epochs {}  rmse {:.4f}  {:.1f} us / Adam step'.format(
        kind, dt, ep, rmse, 1e6 * dt / max(ep) / np.ceil(rows / 200)))
