"""Bounds with the reference's names and methods, evaluated on the GPU."""

from .basic import Ellipsoid, UnitCube, UnitCubeEllipsoidMixture
from .nautilus import NautilusBound
from .neural import NeuralBound
from .union import Union

__all__ = ['UnitCube', 'Ellipsoid', 'UnitCubeEllipsoidMixture', 'Union',
           'NeuralBound', 'NautilusBound']
