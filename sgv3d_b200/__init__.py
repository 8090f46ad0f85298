"""B200-native lift-splat (image -> BEV view transform) for SGV3D / BEVHeight."""
from .shapes import SHAPES, LiftSplatShape, get_shape  # noqa: F401

__version__ = "0.1.0"
