"""``Prior``: named parameters on the unit hypercube.

Mirror of the reference's ``nautilus.prior.Prior`` (nautilus/prior.py:8-181):
same method names, argument meaning and error behaviour.  The arithmetic is
host-side SciPy (arbitrary frozen distributions cannot be kernels; SURVEY.md
section 2.1 row 11), the benchmark configurations use the identity prior.
"""

import numbers

import numpy as np
from scipy.stats import uniform


class Prior:
    """Collects model parameters and maps unit-cube points to their values."""

    def __init__(self):
        self.keys = []
        self.dists = []

    def add_parameter(self, key=None, dist=(0, 1)):
        """Register a parameter (nautilus/prior.py:25-71).

        ``dist``: a number fixes the parameter, a 2-tuple is a uniform range,
        a string ties the parameter to an earlier one, any object with an
        ``isf`` method is used as its distribution.
        """
        if key is None:
            name = 'x_{}'.format(len(self.keys))
        elif not isinstance(key, str):
            raise TypeError("Keyword argument 'key' must be a string.")
        elif key in self.keys:
            raise ValueError("Key '{}' already in key list.".format(key))
        else:
            name = key

        if isinstance(dist, tuple):
            lo, hi = dist
            entry = uniform(loc=lo, scale=hi - lo)
        elif isinstance(dist, numbers.Number) or hasattr(dist, 'isf'):
            entry = dist
        elif isinstance(dist, str):
            if dist not in self.keys or dist == str(key):
                raise ValueError('Key {} not defined previously.'.format(dist))
            entry = dist
            while isinstance(self.dists[self.keys.index(entry)], str):
                entry = self.dists[self.keys.index(entry)]
        else:
            raise TypeError("Keyword argument 'dist' does not have the "
                            "correct type")
        self.keys.append(name)
        self.dists.append(entry)

    def _free(self):
        return [d for d in self.dists if hasattr(d, 'isf')]

    def dimensionality(self):
        """Number of free parameters (nautilus/prior.py:73-83)."""
        return len(self._free())

    def _check(self, points):
        if points.shape[-1] != self.dimensionality():
            raise ValueError('Dimensionality of points does not match prior.')

    def unit_to_physical(self, points):
        """``dist.isf(1 - u)`` per free parameter (nautilus/prior.py:85-120)."""
        points = np.asarray(points)
        self._check(points)
        out = np.zeros_like(points)
        for i, dist in enumerate(self._free()):
            out[..., i] = dist.isf(1 - points[..., i])
        return out

    def physical_to_dictionary(self, phys_points):
        """Pack physical points by key, adding fixed and tied parameters
        (nautilus/prior.py:122-162)."""
        phys_points = np.asarray(phys_points)
        self._check(phys_points)
        result = {}
        column = 0
        for key, dist in zip(self.keys, self.dists):
            if hasattr(dist, 'isf'):
                result[key] = phys_points[..., column]
                column += 1
            elif isinstance(dist, numbers.Number):
                result[key] = np.ones(phys_points[..., 0].shape) * dist
        for key, dist in zip(self.keys, self.dists):
            if isinstance(dist, str):
                result[key] = result[dist]
        return result

    def unit_to_dictionary(self, points):
        """nautilus/prior.py:164-181."""
        return self.physical_to_dictionary(self.unit_to_physical(points))

    # -- the same transforms on CUDA tensors (for TorchLikelihood) ----------
    def unit_to_physical_device(self, points):
        """``unit_to_physical`` on a CUDA tensor f64[n, d]: the inverse
        survival function ``isf(1 - u)`` of every free parameter evaluated
        with torch on the device (nautilus/prior.py:85-120).  Supported
        frozen SciPy distributions: uniform, norm, expon, lognorm (shape s),
        truncnorm is not; anything else raises ``NotImplementedError`` -- use
        a host likelihood for those."""
        import torch
        if points.shape[-1] != self.dimensionality():
            raise ValueError('Dimensionality of points does not match prior.')
        out = torch.empty_like(points)
        for i, dist in enumerate(self._free()):
            out[..., i] = _ppf_device(dist, points[..., i])
        return out

    def physical_to_dictionary_device(self, phys_points):
        import torch
        result = {}
        column = 0
        for key, dist in zip(self.keys, self.dists):
            if hasattr(dist, 'isf'):
                result[key] = phys_points[..., column]
                column += 1
            elif isinstance(dist, numbers.Number):
                result[key] = torch.full_like(phys_points[..., 0], float(dist))
        for key, dist in zip(self.keys, self.dists):
            if isinstance(dist, str):
                result[key] = result[dist]
        return result

    def unit_to_dictionary_device(self, points):
        return self.physical_to_dictionary_device(
            self.unit_to_physical_device(points))


def _ppf_device(dist, u):
    """ppf(u) == isf(1 - u) of a frozen SciPy distribution on a CUDA tensor."""
    import torch
    name = getattr(getattr(dist, 'dist', None), 'name', None)
    args = tuple(getattr(dist, 'args', ()))
    kwds = dict(getattr(dist, 'kwds', {}))
    n_shape = {'uniform': 0, 'norm': 0, 'expon': 0, 'lognorm': 1}.get(name)
    if n_shape is None:
        raise NotImplementedError(
            'no device transform for the prior distribution {!r}; supported: '
            'uniform, norm, expon, lognorm'.format(name))
    shapes, rest = args[:n_shape], args[n_shape:]
    loc = float(kwds.get('loc', rest[0] if len(rest) > 0 else 0.0))
    scale = float(kwds.get('scale', rest[1] if len(rest) > 1 else 1.0))
    if name == 'uniform':
        z = u
    elif name == 'norm':
        z = torch.special.ndtri(u)
    elif name == 'expon':
        z = -torch.log1p(-u)
    else:                                   # lognorm, shape s
        sh = float(kwds.get('s', shapes[0] if shapes else 1.0))
        z = torch.exp(sh * torch.special.ndtri(u))
    return loc + scale * z
