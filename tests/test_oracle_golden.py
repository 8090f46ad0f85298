"""CPU: the oracle (torch port, numpy explicit-order restatement, C restatement) replays the
golden fixtures produced by the reference's own code (tests/golden/make_golden.py) bit-exactly."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import lift_splat_oracle as O
from tests.helpers import frustum_axes, golden_mats, golden_names, kept_mask_np, load_golden, oracle_frustum


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", golden_names())
def test_buffers_and_frustum(name):
    g = load_golden(name)
    s = g["shape"]
    fr = oracle_frustum(s)
    u, v, z = frustum_axes(fr)
    assert np.array_equal(u.numpy(), g["frustum_u"])
    assert np.array_equal(v.numpy(), g["frustum_v"])
    assert np.array_equal(z.numpy(), g["frustum_z"])
    vs, vc, vn = O.grid_buffers(s.x_bound, s.y_bound, s.z_bound)
    assert np.array_equal(vs.numpy(), g["voxel_size"])
    assert np.array_equal(vc.numpy(), g["voxel_coord"])
    assert np.array_equal(vn.numpy(), g["voxel_num"])
    assert tuple(int(t) for t in vn) == s.grid


@pytest.mark.parametrize("name", golden_names())
def test_geometry_and_indices_bit_exact(name):
    g = load_golden(name)
    s = g["shape"]
    m = golden_mats(g)
    fr = oracle_frustum(s)
    vs, vc, vn = O.grid_buffers(s.x_bound, s.y_bound, s.z_bound)
    # 1. torch port, same calls as the reference
    geom = O.geometry_matmul(fr, m["sensor2ego"], m["sensor2virtual"], m["intrin"], m["ida"],
                             m["reference_heights"], m["bda"]).contiguous()
    idx = O.quantize(geom, vc, vs)
    assert _sha(geom.numpy()) == str(g["geom_sha256"])
    assert _sha(idx.numpy()) == str(g["idx_sha256"])
    # 2. explicit-order numpy and C restatements
    ida_inv, mv, me = O.camera_matrices(m["sensor2ego"], m["sensor2virtual"], m["intrin"], m["ida"])
    u, v, z = (t.numpy() for t in frustum_axes(fr))
    bda = m["bda"].numpy() if m["bda"] is not None else None
    rh = m["reference_heights"].numpy()
    geom_c = CO.geometry(CO.ARITH_SEQ, u, v, z, ida_inv.numpy(), mv.numpy(), me.numpy(), rh, bda)
    assert _sha(geom_c) == str(g["geom_sha256"])
    if geom_c.size <= 400_000:
        geom_np = O.geometry_explicit(u, v, z, ida_inv.numpy(), mv.numpy(), me.numpy(), rh, bda)
        assert _sha(geom_np) == str(g["geom_sha256"])
    lower = (vc - vs / 2.0).numpy()
    idx_c = CO.quantize(geom_c, lower, vs.numpy())
    assert _sha(idx_c) == str(g["idx_sha256"])
    assert np.array_equal(O.quantize_np(geom_c, vc.numpy(), vs.numpy()), idx_c)
    assert int(kept_mask_np(idx_c, s.grid).sum()) == int(g["kept_count"])
    if "geom" in g:
        assert np.array_equal(geom_c.view(np.int32), g["geom"].view(np.int32))
        assert np.array_equal(idx_c, g["idx"])


@pytest.mark.parametrize("name", [n for n in golden_names() if "hash" not in n and n != "small_identity"])
def test_voxel_pooling_forward_backward(name):
    """reference VoxelPooling Function (python, naive per-point ext stand-in) vs the oracle's
    index_add_ port, the C restatement and the fused double-precision restatement."""
    g = load_golden(name)
    s = g["shape"]
    X, Y, Z = s.grid
    B, Nc = g["batch"], g["num_cams"]
    idx = torch.from_numpy(g["idx"])
    height, ctx = torch.from_numpy(g["height"]), torch.from_numpy(g["ctx"])
    feat = O.lift(height, ctx)
    feat = feat.reshape(B, Nc, *feat.shape[1:]).permute(0, 1, 3, 4, 5, 2).contiguous()
    bev, pos = O.voxel_pooling_forward(idx, feat, torch.tensor([X, Y, Z]))
    np.testing.assert_allclose(bev.numpy(), g["bev"], rtol=1e-5, atol=1e-6)
    out_c, pos_c = CO.voxel_pooling_forward(g["idx"], feat.numpy(), X, Y, Z)
    np.testing.assert_allclose(out_c.transpose(0, 3, 1, 2), g["bev"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(pos_c, pos.numpy())
    bev64 = CO.lift_splat_forward64(g["idx"], g["height"], g["ctx"], X, Y, Z)
    np.testing.assert_allclose(bev64, g["bev"], rtol=1e-5, atol=1e-6)
    # backward: exact copies, so bitwise
    gref = g["grad_feat"].reshape(B, -1, s.channels)
    gp = O.voxel_pooling_backward(torch.from_numpy(g["grad_bev"]), pos, s.channels)
    assert np.array_equal(gp.numpy(), gref)
    gc = CO.voxel_pooling_backward(g["grad_bev"], pos_c, s.channels)
    assert np.array_equal(gc, gref)
    # fused backward restatement == autograd of the outer product applied to grad_feat
    gh64, gc64 = CO.lift_splat_backward64(g["idx"], g["height"], g["ctx"], g["grad_bev"], X, Y, Z)
    gf = torch.from_numpy(g["grad_feat"]).double()              # (B,Nc,D,fH,fW,C)
    gf = gf.permute(0, 1, 5, 2, 3, 4).reshape(B * Nc, s.channels, s.D, s.fH, s.fW)
    np.testing.assert_allclose(gh64, (gf * ctx.double().unsqueeze(2)).sum(1).numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gc64, (gf * height.double().unsqueeze(1)).sum(2).numpy(), rtol=1e-9, atol=1e-12)


def test_quantize_edge_semantics():
    """Truncation toward zero (not floor), NaN -> 0, saturation: SURVEY.md §7 hard part 2."""
    lower = np.array([0.0, -51.2, -5.0], np.float32)
    size = np.array([0.8, 0.8, 8.0], np.float32)
    pts = np.array([[-0.5, -51.5, -9.0],      # (-1,0) voxel units -> 0, kept
                    [np.nan, np.inf, -np.inf],
                    [1e30, -1e30, 2.99],
                    [0.8, -50.0, 3.0]], np.float32)
    idx = CO.quantize(pts, lower, size)
    assert idx.tolist() == [[0, 0, 0], [0, 2147483647, -2147483648],
                            [2147483647, -2147483648, 0], [1, 1, 1]]
    vc = lower + size / np.float32(2.0)
    assert np.array_equal(O.quantize_np(pts, vc, size), idx)


@pytest.mark.parametrize("name", golden_names())
def test_product_build_frustum_matches_the_reference_buffer(name):
    """sgv3d_b200.build_frustum (the PRODUCT's copy of LSSFPN.create_frustum, lss_fpn.py:325-348) against the frustum
    buffer the reference's own code produced (tests/golden/make_golden.py): bit-exact u / v / z axes."""
    from sgv3d_b200.view_transform import build_frustum
    g = load_golden(name)
    shape = g["shape"]
    fr = build_frustum(shape.final_dim, shape.downsample, shape.d_bound)
    u, v, z = (t.numpy() for t in frustum_axes(fr))
    for got, key in ((u, "frustum_u"), (v, "frustum_v"), (z, "frustum_z")):
        want = g[key]
        assert got.shape == want.shape and (got.view(np.int32) == want.view(np.int32)).all(), key
    assert float(fr[..., 3].min()) == 1.0 == float(fr[..., 3].max())
