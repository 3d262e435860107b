"""Neural-network emulator of the likelihood score.

Host-side mirror of ``nautilus/neural.py``.  ``train`` fits the ensemble with
the persistent on-chip trainer (csrc/nb200_mlp_fit.cu), ``predict`` runs the
forward pass on the GPU (fp64 parity kernel or the tcgen05 tf32 kernel).  The
fitted parameters are kept in the same shape scikit-learn exposes
(``coefs_``, ``intercepts_``, ``n_layers_``), which is what the reference
serialises (nautilus/neural.py:139-143) and what ``n_net`` inspects
(nautilus/bounds/nautilus.py:286-288).
"""

import warnings

import numpy as np
import torch

from . import ops
from ._device import default_device, to_device

DEFAULT_KWARGS = dict(hidden_layer_sizes=(100, 50, 20), alpha=0,
                      learning_rate_init=1e-2, max_iter=10000, tol=0,
                      n_iter_no_change=10)


class FittedNetwork:
    """The attributes of a fitted ``MLPRegressor`` the rest of the code reads.
    """

    def __init__(self, coefs, intercepts, n_iter, loss):
        self.coefs_ = coefs
        self.intercepts_ = intercepts
        self.n_layers_ = len(coefs) + 1
        self.n_iter_ = int(n_iter)
        self.loss_ = float(loss)


class NeuralNetworkEmulator:
    """Ensemble of small MLPs on standardised inputs
    (nautilus/neural.py:35-187)."""

    mode = ops.MLP_F64       # arithmetic of predict(); class-wide default

    @classmethod
    def train(cls, x, y, n_networks=4, neural_network_kwargs={}, pool=None,
              seed=0):
        """Standardise ``x`` and fit ``n_networks`` networks
        (nautilus/neural.py:50-98).  ``pool`` is accepted for API
        compatibility; the networks train concurrently, one CTA cluster
        each."""
        return cls.train_async(x, y, n_networks, neural_network_kwargs, pool,
                               seed).wait()

    @classmethod
    def train_async(cls, x, y, n_networks=4, neural_network_kwargs={},
                    pool=None, seed=0):
        """``train`` without the final synchronisation: the fit kernel is
        enqueued on the current CUDA stream and ``wait()`` collects the
        weights.  Lets the caller build the rest of a bound while the
        networks train (bounds/nautilus.py: the outer union does not depend
        on them)."""
        emulator = cls()
        x = np.asarray(x, dtype=float)
        y = np.asarray(y, dtype=float)
        emulator.mean = np.mean(x, axis=0)
        emulator.scale = np.std(x, axis=0)

        kwargs = dict(DEFAULT_KWARGS)
        kwargs.update(neural_network_kwargs)
        if 'random_state' in kwargs:
            warnings.warn("The 'random_state' keyword argument passed to the"
                          " neural network is ignored.", Warning, stacklevel=2)
            del kwargs['random_state']
        supported = set(DEFAULT_KWARGS) | {'batch_size', 'beta_1', 'beta_2',
                                           'epsilon'}
        unknown = set(kwargs) - supported
        if unknown:
            raise ValueError('Unsupported neural network arguments: {}'
                             .format(sorted(unknown)))
        if kwargs['alpha'] != 0:
            raise ValueError("L2 penalty 'alpha' is not supported.")

        hidden = tuple(np.atleast_1d(kwargs['hidden_layer_sizes']))
        sizes = (x.shape[1], ) + tuple(int(h) for h in hidden) + (1, )
        batch_size = kwargs.get('batch_size', 'auto')
        if batch_size == 'auto':
            batch_size = min(200, len(x))
        dev = default_device()
        # The callers hand the rows over sorted by likelihood
        # (Sampler.add_bound sorts before NautilusBound.compute) and the
        # trainer walks them in a per-epoch AFFINE order i -> (a i + b) mod M:
        # on rank-ordered rows a minibatch would be an arithmetic progression
        # through the likelihood ranks (a few narrow bands when a / M is close
        # to a simple fraction).  One seeded Fisher-Yates shuffle of the rows
        # removes the ordering; an arithmetic progression through shuffled
        # rows is an unstructured subset, like sklearn's per-epoch shuffle
        # (_multilayer_perceptron.py:708).
        order = np.random.default_rng(
            [int(seed) & 0xFFFFFFFFFFFFFFFF, 0x5AFF1E]).permutation(len(x))
        xs = torch.from_numpy(
            ((x - emulator.mean) / emulator.scale)[order]).to(dev)
        y = y[order]
        params, n_iter, loss = ops.mlp_fit(
            xs.contiguous(), torch.from_numpy(y).to(dev), sizes, n_networks,
            seed=seed, lr=kwargs['learning_rate_init'],
            beta1=kwargs.get('beta_1', 0.9), beta2=kwargs.get('beta_2', 0.999),
            eps=kwargs.get('epsilon', 1e-8), batch_size=batch_size,
            max_epochs=kwargs['max_iter'], tol=kwargs['tol'],
            patience=kwargs['n_iter_no_change'])
        emulator.n_train_ = len(x)      # rows the ensemble was fitted on
        emulator._pending = (params, n_iter, loss, sizes, n_networks,
                             torch.cuda.current_stream())
        return emulator

    def wait(self):
        """Collect the result of ``train_async`` (idempotent)."""
        pending = getattr(self, '_pending', None)
        if pending is None:
            return self
        params, n_iter, loss, sizes, n_networks, stream = pending
        self._pending = None
        emulator = self
        stream.synchronize()
        params = params.cpu().numpy()
        n_iter = n_iter.cpu().numpy()
        loss = loss.cpu().numpy()
        emulator.neural_networks = []
        for i in range(n_networks):
            coefs, intercepts, off = [], [], 0
            for fi, fo in zip(sizes[:-1], sizes[1:]):
                coefs.append(params[i, off:off + fi * fo].reshape(fi, fo))
                off += fi * fo
                intercepts.append(params[i, off:off + fo].copy())
                off += fo
            emulator.neural_networks.append(
                FittedNetwork(coefs, intercepts, n_iter[i], loss[i]))
        emulator._stack = None
        return emulator

    def write(self, group):
        """Layout of nautilus/neural.py:118-146: per network i the scalar
        attributes as '<name>_<i>' and the layers as 'coefs_<k>_<i>' /
        'intercepts_<k>_<i>'.  The few extra attributes scikit-learn's
        ``predict`` needs are written too, so that the reference can load the
        emulator into an ``MLPRegressor``."""
        self.wait()
        group.attrs['n_networks'] = len(self.neural_networks)
        for i, network in enumerate(self.neural_networks):
            attrs = dict(n_layers_=network.n_layers_, n_iter_=network.n_iter_,
                         loss_=network.loss_, activation='relu',
                         out_activation_='identity', n_outputs_=1,
                         n_features_in_=network.coefs_[0].shape[0])
            for key, value in attrs.items():
                group.attrs['{}_{}'.format(key, i)] = value
            for k in range(network.n_layers_ - 1):
                group.create_dataset('coefs_{}_{}'.format(k, i),
                                     data=network.coefs_[k])
                group.create_dataset('intercepts_{}_{}'.format(k, i),
                                     data=network.intercepts_[k])
        group.create_dataset('mean', data=self.mean)
        group.create_dataset('scale', data=self.scale)

    @classmethod
    def read(cls, group):
        """(nautilus/neural.py:148-187)."""
        emulator = cls()
        emulator.mean = np.array(group['mean'], dtype=float)
        emulator.scale = np.array(group['scale'], dtype=float)
        emulator.neural_networks = []
        for i in range(int(group.attrs['n_networks'])):
            coefs, intercepts = [], []
            while 'coefs_{}_{}'.format(len(coefs), i) in group:
                k = len(coefs)
                coefs.append(np.array(group['coefs_{}_{}'.format(k, i)],
                                      dtype=float))
                intercepts.append(np.array(
                    group['intercepts_{}_{}'.format(k, i)], dtype=float))
            stored = {key: group.attrs['{}_{}'.format(key, i)]
                      if '{}_{}'.format(key, i) in group.attrs else default
                      for key, default in (('n_iter_', 0), ('loss_', np.nan))}
            emulator.neural_networks.append(FittedNetwork(
                coefs, intercepts, stored['n_iter_'], stored['loss_']))
        emulator._stack = None
        return emulator

    def emu_spec(self):
        return dict(mean=self.mean, scale=self.scale,
                    coefs=[n.coefs_ for n in self.neural_networks],
                    intercepts=[n.intercepts_ for n in self.neural_networks])

    def _device_stack(self):
        if getattr(self, '_stack', None) is None:
            d = len(self.mean)
            ident = dict(c=np.zeros(d), B=np.eye(d), B_inv=np.eye(d))
            spec = dict(kind='nautilus', n_dim=d, unit=False,
                        log_v_all=np.zeros(1),
                        mixtures=[dict(dim_cube=np.zeros(d, bool), ell=ident)],
                        neural=[dict(ell=ident, emulator=self.emu_spec(),
                                     score_predict_min=0.0)])
            self._stack = ops.DeviceStack([spec], device=default_device())
        return self._stack

    def predict(self, x, mode=None):
        """Mean network output on ``(x - mean) / scale``
        (nautilus/neural.py:100-116)."""
        t, restore = to_device(x, len(self.mean))
        mode = self.mode if mode is None else mode
        return restore(self._device_stack().mlp_predict(0, 0, t, mode=mode))
