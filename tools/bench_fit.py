"""Time the on-chip emulator trainer against scikit-learn on the config-2
training problem: ~4 200 rows x 30-D, 4 networks."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))


def main():
    import torch
    from nautilus_b200.neural import NeuralNetworkEmulator
    rng = np.random.default_rng(0)
    d, m = 30, 4200
    x = rng.normal(size=(m, d)) * 0.4
    r = np.linalg.norm(x, axis=1)
    y = np.argsort(np.argsort(-r)) / m
    NeuralNetworkEmulator.train(x[:500], y[:500], n_networks=1)   # warm up
    torch.cuda.synchronize()
    t0 = time.time()
    emu = NeuralNetworkEmulator.train(x, y, n_networks=4)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    pred = emu.predict(x)
    out = dict(rows=m, n_dim=d, n_networks=4, gpu_fit_s=t_gpu,
               gpu_epochs=[n.n_iter_ for n in emu.neural_networks],
               gpu_rmse_over_std=float(np.sqrt(np.mean((pred - y)**2)) /
                                       np.std(y)))
    if '--sklearn' in sys.argv:
        from sklearn.neural_network import MLPRegressor
        from threadpoolctl import threadpool_limits
        xs = (x - x.mean(0)) / x.std(0)
        t0 = time.time()
        with threadpool_limits(limits=1):
            nets = [MLPRegressor(hidden_layer_sizes=(100, 50, 20), alpha=0,
                                 learning_rate_init=1e-2, max_iter=10000,
                                 tol=0, n_iter_no_change=10,
                                 random_state=i).fit(xs, y) for i in range(4)]
        out['sklearn_fit_s_1core'] = time.time() - t0
        out['sklearn_epochs'] = [n.n_iter_ for n in nets]
        p = np.mean([n.predict(xs) for n in nets], axis=0)
        out['sklearn_rmse_over_std'] = float(
            np.sqrt(np.mean((p - y)**2)) / np.std(y))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
