"""CPU: the C-ABI library loads, exports every symbol include/sgv3d_b200.h declares, and its
argument validation / error reporting works without touching a GPU."""
import os
import re

import pytest

from sgv3d_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sgv3d_b200.h")).read()
    declared = set(re.findall(r"SGV3D_API\s+[\w\s\*]+?\b(sgv3d_\w+)\s*\(", header))
    assert declared == set(N.SYMBOLS), declared ^ set(N.SYMBOLS)
    L = N.lib()
    for name in declared:
        assert hasattr(L, name)
    assert L.sgv3d_abi_version() == N.ABI_VERSION


def test_workspace_queries_and_argument_validation():
    L = N.lib()
    d = N.LiftSplatDesc(B=8, Nc=1, D=90, fH=54, fW=96, C=80, X=128, Y=128, Z=1, arith=0, ctx_dtype=0)
    assert L.sgv3d_lift_splat_workspace_bytes(d) > 0
    assert L.sgv3d_voxel_pooling_workspace_bytes(8, 466560, 80, 128, 128, 1) > 0
    assert L.sgv3d_voxel_pooling_backward_workspace_bytes(8, 80, 128, 128) == 8 * 80 * 128 * 128 * 4
    bad = N.LiftSplatDesc(B=1, Nc=1, D=90, fH=54, fW=96, C=80, X=1024, Y=1024, Z=1, arith=0, ctx_dtype=0)
    assert L.sgv3d_lift_splat_workspace_bytes(bad) == 0
    assert b"voxels per frame" in L.sgv3d_last_error()
    lo, sz = N.host_f32x3([0, -51.2, -5]), N.host_f32x3([0.8, 0.8, 8])
    rc = L.sgv3d_geometry_quantize(7, 1, 1, 4, 3, 5, 1, 1, 1, 1, 1, 1, 0, 1, lo, sz, 0, 0, 0)
    assert rc == 1 and b"bad arith" in L.sgv3d_last_error()
    with pytest.raises(RuntimeError, match="bad arith"):
        N.check(rc)
    # B == 0 is a no-op that must not touch the device
    assert L.sgv3d_geometry_quantize(0, 0, 1, 4, 3, 5, 1, 1, 1, 1, 1, 1, 0, 1, lo, sz, 0, 0, 0) == 0
    # the 4x4 prep entry points validate before they launch
    assert L.sgv3d_inverse4x4(-1, 0, 0, 0, 0, 0, 0, 0) == 1 and b"bad count" in L.sgv3d_last_error()
    assert L.sgv3d_inverse4x4(0, 0, 0, 0, 0, 0, 0, 0) == 0
    assert L.sgv3d_inverse4x4(4, 16, 0, 32, 48, 0, 64, 0) == 1 and b"pair up" in L.sgv3d_last_error()
    assert L.sgv3d_inverse4x4(4, 16, 0, 0, 40, 0, 0, 0) == 1 and b"aligned" in L.sgv3d_last_error()
    assert L.sgv3d_camera_prep(4, N.ARITH_PAIR, 16, 32, 48, 64, 80, 96, 112, 0) == 1
    assert b"SEQ or FMA" in L.sgv3d_last_error()
    assert L.sgv3d_camera_prep(0, N.ARITH_FMA, 0, 0, 0, 0, 0, 0, 0, 0) == 0
    d2 = N.LiftSplatDesc(B=1, Nc=1, D=90, fH=54, fW=96, C=87, X=128, Y=128, Z=1, arith=0, ctx_dtype=0)
    assert L.sgv3d_lift_splat_forward_bsm(d2, 16, 16, 16, 87, 0, 0.45, 16, 16, 0, 0) == 1
    assert b"semantic_channels" in L.sgv3d_last_error()
    # too-small workspace is reported, not written past
    rc = L.sgv3d_voxel_pooling_forward(1, 100, 8, 4, 4, 1, 1, 1, 1, 0, 1, 16, 0)
    assert rc == 2 and b"workspace" in L.sgv3d_last_error()


def test_cpu_tensors_are_rejected_not_silently_computed():
    import torch
    from sgv3d_b200 import voxel_pooling
    with pytest.raises(RuntimeError, match="CUDA"):
        voxel_pooling(torch.zeros(1, 4, 3, dtype=torch.int32), torch.zeros(1, 4, 2), [2, 2, 1])
    from sgv3d_b200.view_transform import LiftSplatPlan, build_frustum
    fr = build_frustum((48, 80), 16, (-2.0, 0.0, 4))
    eye = torch.eye(4).view(1, 1, 4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        LiftSplatPlan(fr, eye, eye, eye, eye, torch.ones(1, 1), None, torch.zeros(3), torch.ones(3), (4, 4, 1), 8)
