"""B200-native lift-splat (image -> BEV view transform) for SGV3D / BEVHeight.

Public surface (mirrors the reference's operator / call-site interface for this path):

* ``sgv3d_b200.ops.voxel_pooling.voxel_pooling(geom_xyz, input_features, voxel_num)``
      drop-in for ``ops.voxel_pooling.voxel_pooling``           (ops/voxel_pooling/voxel_pooling.py:72)
* ``LiftSplat`` / ``lift_splat`` / ``LiftSplatPlan``
      fused replacement of the lift-splat block of ``_forward_single_sweep``
                                                                (layers/backbones/lss_fpn.py:462-495)
* ``geometry_indices``
      ``get_geometry`` + quantisation                           (layers/backbones/lss_fpn.py:372-401,487-488)

All compute runs in ``csrc/libsgv3d_b200.so`` (hand-written sm_100a CUDA behind the C ABI declared in
``include/sgv3d_b200.h``).  There is no CPU / eager fallback: calling any entry point without the
library, or with CPU tensors, raises.
"""
from .shapes import SHAPES, LiftSplatShape, get_shape  # noqa: F401
from .ops.voxel_pooling import VoxelPooling, voxel_pooling  # noqa: F401
from .view_transform import (LiftSplat, LiftSplatGraph, LiftSplatPlan, build_frustum,  # noqa: F401
                             camera_matrices, geometry_indices, lift_splat)

__version__ = "0.1.0"
