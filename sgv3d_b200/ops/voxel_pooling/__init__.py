from .voxel_pooling import VoxelPooling, voxel_pooling

__all__ = ["voxel_pooling", "VoxelPooling"]
