"""Time the device enclosing-ellipsoid iteration (k_mvee) against the NumPy
restatement on the config-2 live set (2 000 points, 30-D)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))


def main():
    import torch
    from nautilus_b200.bounds import _construct
    rng = np.random.default_rng(0)
    out = []
    for n, d in [(2000, 30), (4000, 50), (500, 10)]:
        z = rng.normal(size=(n, d))
        z *= (rng.uniform(size=(n, 1))**(1.0 / d) /
              np.linalg.norm(z, axis=1)[:, None])
        x = 0.5 + 0.1 * z @ (np.eye(d) + 0.3 * rng.normal(size=(d, d))).T
        _construct._MVEE_CACHE.clear()
        _construct.enclosing_ellipsoid(x, device='cuda')      # warm up
        torch.cuda.synchronize()
        t_dev = []
        for _ in range(3):
            _construct._MVEE_CACHE.clear()
            t0 = time.time()
            c1, a1, _ = _construct.enclosing_ellipsoid(x, device='cuda')
            t_dev.append(time.time() - t0)
        _construct._MVEE_CACHE.clear()
        t0 = time.time()
        c0, a0, _ = _construct.enclosing_ellipsoid(x)
        t_host = time.time() - t0
        out.append(dict(n=n, d=d, device_s=min(t_dev), host_s=t_host,
                        max_dc=float(np.max(np.abs(c0 - c1))),
                        rel_da=float(np.max(np.abs(a0 - a1)) /
                                     np.max(np.abs(a0)))))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
