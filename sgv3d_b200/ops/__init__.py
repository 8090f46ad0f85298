"""Operator package, laid out like the reference's ``ops/`` (ops/voxel_pooling/__init__.py:1-3)."""
