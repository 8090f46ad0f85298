"""GPU: sgv3d_geometry_quantize vs the oracle -- voxel indices must be BIT-EXACT."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import c_oracle as CO
from oracle import lift_splat_oracle as O
from sgv3d_b200 import get_shape
from sgv3d_b200.synthetic import make_mats
from tests.helpers import frustum_axes, golden_mats, golden_names, load_golden, oracle_frustum

pytestmark = pytest.mark.gpu


def _run_kernel(shape, mats_cpu, arith):
    from sgv3d_b200.view_transform import geometry_indices
    fr = oracle_frustum(shape)
    vs, vc, _ = O.grid_buffers(shape.x_bound, shape.y_bound, shape.z_bound)
    m = {k: (v.cuda() if v is not None else None) for k, v in mats_cpu.items()}
    # per-camera 4x4 prep on the CPU tensors so that kernel and oracle see identical operands
    idx, xyz = geometry_indices(fr, m["sensor2ego"], m["sensor2virtual"], m["intrin"], m["ida"],
                                m["reference_heights"], m["bda"], vc, vs, arith=arith, return_xyz=True)
    return idx.cpu().numpy(), xyz.cpu().numpy()


def _oracle(shape, mats_cpu, mode, device_mats=None):
    fr = oracle_frustum(shape)
    vs, vc, _ = O.grid_buffers(shape.x_bound, shape.y_bound, shape.z_bound)
    src = device_mats or mats_cpu
    ida_inv, mv, me = O.camera_matrices(src["sensor2ego"], src["sensor2virtual"], src["intrin"], src["ida"])
    u, v, z = (t.numpy() for t in frustum_axes(fr))
    bda = mats_cpu["bda"].numpy() if mats_cpu["bda"] is not None else None
    xyz = CO.geometry(mode, u, v, z, ida_inv.cpu().numpy(), mv.cpu().numpy(), me.cpu().numpy(),
                      mats_cpu["reference_heights"].numpy(), bda)
    idx = CO.quantize(xyz, (vc - vs / 2.0).numpy(), vs.numpy())
    return idx, xyz


def _mats(shape, batch, num_cams, seed, bda):
    m = make_mats(shape, batch, num_cams, seed=seed, bda=bda)
    return {"sensor2ego": m["sensor2ego"], "sensor2virtual": m["sensor2virtual"], "intrin": m["intrin"],
            "ida": m["ida"], "reference_heights": m["reference_heights"], "bda": m["bda"]}


@pytest.mark.parametrize("arith,mode", [(0, CO.ARITH_SEQ), (1, CO.ARITH_FMA), (2, CO.ARITH_PAIR)])
@pytest.mark.parametrize("shape_name,batch,num_cams,bda", [
    ("tiny", 2, 2, None), ("small", 3, 1, "random"), ("dair_r50", 2, 1, "identity"),
    ("rope3d_r50", 1, 1, "random"), ("rope3d_native", 1, 1, "identity"), ("sgv3d_bsm_r50", 1, 1, "identity"),
])
def test_indices_bit_exact_vs_c_oracle(shape_name, batch, num_cams, bda, arith, mode):
    shape = get_shape(shape_name)
    mats = _mats(shape, batch, num_cams, seed=21, bda=bda)
    # the 4x4 prep runs on the GPU inside geometry_indices; feed the oracle the same 16-float operands
    dev = {k: (v.cuda() if v is not None else None) for k, v in mats.items()}
    idx_o, xyz_o = _oracle(shape, mats, mode, device_mats=dev)
    idx_k, xyz_k = _run_kernel(shape, mats, arith)
    bad = int((idx_k != idx_o).any(-1).sum())
    assert bad == 0, f"{bad} of {idx_o.size // 3} points differ"
    assert np.array_equal(xyz_k.view(np.int32), xyz_o.view(np.int32))


@pytest.mark.parametrize("name", golden_names())
def test_golden_indices_seq(name):
    """Reference-generated fixtures (reference run on CPU): SEQ arithmetic reproduces them exactly,
    given the same 4x4 operands (the per-camera prep is done on CPU here, as the reference did)."""
    from sgv3d_b200 import _native as N
    g = load_golden(name)
    s = g["shape"]
    m = golden_mats(g)
    ida_inv, mv, me = O.camera_matrices(m["sensor2ego"], m["sensor2virtual"], m["intrin"], m["ida"])
    B, Nc = g["batch"], g["num_cams"]
    idx = torch.empty(B, Nc, s.D, s.fH, s.fW, 3, dtype=torch.int32, device="cuda")
    dev = lambda a: torch.as_tensor(a).float().contiguous().cuda()
    u, v, z = dev(g["frustum_u"]), dev(g["frustum_v"]), dev(g["frustum_z"])
    a, b_, c = dev(ida_inv), dev(mv), dev(me)
    rh = dev(m["reference_heights"]).reshape(-1)
    bda = dev(m["bda"]) if m["bda"] is not None else None
    vc, vs = torch.from_numpy(g["voxel_coord"]), torch.from_numpy(g["voxel_size"])
    lower, size = N.host_f32x3((vc - vs / 2.0).tolist()), N.host_f32x3(vs.tolist())
    N.check(N.lib().sgv3d_geometry_quantize(N.ARITH_SEQ, B, Nc, s.D, s.fH, s.fW, N.ptr(u), N.ptr(v), N.ptr(z),
                                            N.ptr(a), N.ptr(b_), N.ptr(c), N.ptr(bda), N.ptr(rh), lower, size,
                                            N.ptr(idx), 0, N.current_stream()))
    got = idx.cpu().numpy()
    assert hashlib.sha256(got.tobytes()).hexdigest() == str(g["idx_sha256"])
    if "idx" in g:
        assert np.array_equal(got, g["idx"])


def test_degenerate_rays_follow_gpu_cast_semantics():
    """pv.y == 0 (ray parallel to the ground) -> inf/NaN coordinates; cvt.rzi saturates and maps NaN
    to voxel 0 (kept!) exactly like `.int()` on a CUDA tensor (SURVEY.md §7 hard part 2)."""
    shape = get_shape("small")
    mats = _mats(shape, 1, 1, seed=5, bda="identity")
    mats["sensor2virtual"] = torch.eye(4).view(1, 1, 4, 4).clone()
    k = torch.eye(4)
    k[0, 0] = k[1, 1] = 100.0                        # cx = cy = 0: the v = 0 feature row has pv.y == 0
    mats["intrin"] = k.view(1, 1, 4, 4).clone()
    mats["ida"] = torch.eye(4).view(1, 1, 4, 4).clone()
    dev = {kk: (vv.cuda() if vv is not None else None) for kk, vv in mats.items()}
    for arith, mode in ((0, CO.ARITH_SEQ), (1, CO.ARITH_FMA), (2, CO.ARITH_PAIR)):
        idx_o, _ = _oracle(shape, mats, mode, device_mats=dev)
        idx_k, _ = _run_kernel(shape, mats, arith)
        assert np.array_equal(idx_k, idx_o)
    # the v = 0 feature row: 0 * inf = NaN in every coordinate -> index (0, 0, 0), i.e. KEPT in voxel 0
    assert (idx_o[0, 0, :, 0] == 0).all() and (idx_o[0, 0, :, 1:, :, 0] != 0).any()


@pytest.mark.parametrize("shape_name,batch,bda", [("dair_r50", 2, "identity"), ("rope3d_r50", 2, "random"),
                                                  ("sgv3d_bsm_r50", 1, "identity"), ("small", 4, None)])
def test_default_arith_matches_reference_port_on_the_gpu(shape_name, batch, bda):
    """The device oracle: the torch port of get_geometry (same calls, same broadcast shapes as
    lss_fpn.py:350-401) executed ON THE GPU, i.e. what the reference computes in production
    (cuBLAS bmm).  Default arithmetic (PAIR) must reproduce its indices and coordinates bit for bit."""
    from sgv3d_b200.view_transform import default_arith, geometry_indices
    from sgv3d_b200 import _native as N
    assert default_arith() == N.ARITH_PAIR
    shape = get_shape(shape_name)
    mats = _mats(shape, batch, 1, seed=91, bda=bda)
    dev = {k: (v.cuda() if v is not None else None) for k, v in mats.items()}
    fr = oracle_frustum(shape)
    vs, vc, _ = O.grid_buffers(shape.x_bound, shape.y_bound, shape.z_bound)
    geom = O.geometry_matmul(fr.cuda(), dev["sensor2ego"], dev["sensor2virtual"], dev["intrin"], dev["ida"],
                             dev["reference_heights"], dev["bda"])
    idx_ref = O.quantize(geom, vc.cuda(), vs.cuda())
    idx, xyz = geometry_indices(fr, dev["sensor2ego"], dev["sensor2virtual"], dev["intrin"], dev["ida"],
                                dev["reference_heights"], dev["bda"], vc, vs, return_xyz=True)
    assert int((idx != idx_ref).any(-1).sum()) == 0
    assert torch.equal(xyz.view(torch.int32), geom.contiguous().view(torch.int32))
