"""Synthetic likelihoods of the benchmark configurations.

Each object has two faces that integrate the same function: ``__call__`` is
vectorised NumPy on unit-cube points (what one hands to the reference's
``Sampler(..., vectorized=True)`` for the CPU baseline), and
``(like_id, params)`` is the device functor evaluated inside the cycle
(``k_loglike`` in csrc/nb200_kernels.cu).  tests/test_gpu_parity.py checks the
two against each other.
"""

import numpy as np
from scipy.special import erf, logsumexp

from . import ops


class DeviceLikelihood:
    like_id = -1
    log_z_true = None

    def params(self):
        raise NotImplementedError

    def device_params(self, device):
        import torch
        return torch.from_numpy(np.ascontiguousarray(
            self.params(), dtype=np.float64)).to(device)


class TorchLikelihood(DeviceLikelihood):
    """Plug-in device likelihood: any callable on CUDA tensors.

    ``fn`` receives what the reference hands a vectorised likelihood
    (nautilus/sampler.py:856-873) -- the prior-transformed points as a CUDA
    float64 tensor ``[n, d]`` (or a dict of CUDA tensors per parameter when
    the sampler passes dictionaries) -- and returns ``log L`` as a CUDA
    float64 tensor ``[n]``.  The sampler then keeps the whole batch loop on
    the device: proposal, neural filter and later-bound exclusion in
    ``nb200_cycle`` (no built-in likelihood), ``fn`` on the compacted rows in
    the arena, log-sum-exp / counters by ``nb200_stats``; the host reads the
    row count and 96 bytes of sums per raw batch.  The prior transform runs
    on the device too (``Prior.unit_to_*_device`` or a torch-compatible prior
    callable)."""
    like_id = -1

    def __init__(self, fn, n_dim=None, log_z_true=None):
        self.fn = fn
        self.n_dim = n_dim
        self.log_z_true = log_z_true

    def torch_call(self, args):
        import torch
        out = self.fn(args)
        if not (isinstance(out, torch.Tensor) and out.is_cuda):
            raise TypeError('a TorchLikelihood must return a CUDA tensor')
        return out.to(torch.float64).reshape(-1).contiguous()

    def __call__(self, x):
        """NumPy face (tests, CPU comparisons): upload, evaluate, download."""
        import torch
        t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64),
                            device='cuda')
        return self.torch_call(t).cpu().numpy()

    def params(self):
        return np.zeros(0)


class Gaussian(DeviceLikelihood):
    """N(mu, sigma^2 I), normalised: log Z = sum_i log(Phi-mass in [0,1])."""
    like_id = ops.LIKE_GAUSSIAN

    def __init__(self, n_dim, mu=0.5, sigma=0.1):
        self.n_dim = n_dim
        self.mu = np.broadcast_to(np.asarray(mu, dtype=float), (n_dim,)).copy()
        self.sigma = float(sigma)
        self.inv_sigma2 = 1.0 / self.sigma**2
        self.norm = -0.5 * n_dim * np.log(2 * np.pi * self.sigma**2)
        s = self.sigma * np.sqrt(2)
        self.log_z_true = float(np.sum(np.log(
            0.5 * (erf((1 - self.mu) / s) - erf(-self.mu / s)))))

    def __call__(self, x):
        return (-0.5 * self.inv_sigma2 * np.sum((x - self.mu)**2, axis=-1) +
                self.norm)

    def params(self):
        return np.concatenate([[self.inv_sigma2, self.norm], self.mu])


class Rosenbrock(DeviceLikelihood):
    """log L = -sum_i [100 (t_{i+1} - t_i^2)^2 + (1 - t_i)^2], t = lo + w u."""
    like_id = ops.LIKE_ROSENBROCK

    def __init__(self, n_dim, lo=-5.0, width=10.0):
        self.n_dim = n_dim
        self.lo, self.width = float(lo), float(width)

    def __call__(self, x):
        t = self.lo + self.width * x
        return -np.sum(100.0 * (t[..., 1:] - t[..., :-1]**2)**2 +
                       (1.0 - t[..., :-1])**2, axis=-1)

    def params(self):
        return np.array([self.lo, self.width])


class GaussianMixture(DeviceLikelihood):
    """Equal-weight mixture of M isotropic Gaussians (config 4)."""
    like_id = ops.LIKE_MIXTURE

    def __init__(self, mus, sigma=0.03):
        self.mus = np.atleast_2d(np.asarray(mus, dtype=float))
        self.n_dim = self.mus.shape[1]
        self.sigma = float(sigma)
        self.inv_sigma2 = 1.0 / self.sigma**2
        self.norm = -0.5 * self.n_dim * np.log(2 * np.pi * self.sigma**2)
        self.log_z_true = 0.0

    def __call__(self, x):
        d2 = np.sum((x[..., None, :] - self.mus)**2, axis=-1)
        return (logsumexp(-0.5 * self.inv_sigma2 * d2, axis=-1) -
                np.log(len(self.mus)) + self.norm)

    def params(self):
        return np.concatenate([[len(self.mus), self.inv_sigma2, self.norm],
                               self.mus.ravel()])


class EquicorrelatedGaussian(DeviceLikelihood):
    """N(mu, sigma^2 [(1-rho) I + rho 11^T]) (config 5), closed-form inverse:
    log L = -0.5 (a S2 - b S1^2) + norm with S1 = sum y, S2 = sum y^2."""
    like_id = ops.LIKE_EQUICORR

    def __init__(self, n_dim, mu=0.5, sigma=0.05, rho=0.5):
        self.n_dim = n_dim
        self.mu = np.broadcast_to(np.asarray(mu, dtype=float), (n_dim,)).copy()
        self.sigma, self.rho = float(sigma), float(rho)
        self.a = 1.0 / (sigma**2 * (1 - rho))
        self.b = self.a * rho / (1 + (n_dim - 1) * rho)
        log_det = (n_dim * np.log(sigma**2) + (n_dim - 1) * np.log(1 - rho) +
                   np.log(1 + (n_dim - 1) * rho))
        self.norm = -0.5 * (n_dim * np.log(2 * np.pi) + log_det)
        self.log_z_true = 0.0

    def __call__(self, x):
        y = x - self.mu
        s1 = np.sum(y, axis=-1)
        s2 = np.sum(y * y, axis=-1)
        return -0.5 * (self.a * s2 - self.b * s1 * s1) + self.norm

    def params(self):
        return np.concatenate([[self.a, self.b, self.norm], self.mu])
