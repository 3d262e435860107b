"""nautilus_b200: B200-native inner loop of importance nested sampling.

Same public surface as johannesulf/nautilus (``Prior``, ``Sampler``;
nautilus/__init__.py:3-9).  Importing the package does not need a GPU; using
it does (there is no CPU fallback).
"""

from .prior import Prior

__version__ = '0.1.0'
__all__ = ['Prior', 'Sampler']


def __getattr__(name):
    # Sampler pulls in torch; import lazily so that host-only tools (packing,
    # the ABI check) stay light
    if name == 'Sampler':
        from .sampler import Sampler
        return Sampler
    raise AttributeError(name)
