#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by executing the REFERENCE's own code.

Runs only in the build container (needs /root/reference, read-only).  It imports
``layers/backbones/lss_fpn.py`` and ``ops/voxel_pooling/voxel_pooling.py`` UNMODIFIED, with
``sys.modules`` stubs for the third-party packages that are not installed (mmcv, mmdet,
mmdet3d -- SURVEY.md Appendix D) and, for the CUDA-only extension module, a naive pure-python
stand-in for ``voxel_pooling_forward_wrapper`` that follows
ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:16-34 literally (one point at a time).

Outputs (committed): golden_<case>.npz holding the inputs and the reference's outputs.
    python tests/golden/make_golden.py
"""
import ast
import hashlib
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from sgv3d_b200.shapes import get_shape  # noqa: E402
from sgv3d_b200.synthetic import make_mats  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _naive_ext_forward(b, n, c, nx, ny, nz, geom, feat, out, pos):
    nx, ny, nz = int(nx), int(ny), int(nz)
    for pt in range(b * n):
        bi = pt // n
        x, y, z = (int(v) for v in geom.view(-1, 3)[pt])
        if x < 0 or x >= nx or y < 0 or y >= ny or z < 0 or z >= nz:
            continue
        pos.view(-1, 3)[pt] = torch.tensor([bi, y, x], dtype=pos.dtype)
        out[bi, y, x] += feat.view(-1, c)[pt]
    return 1


def load_reference():
    _stub("mmcv"); _stub("mmcv.cnn", build_conv_layer=None)
    _stub("mmdet3d"); _stub("mmdet3d.models", build_neck=None)
    _stub("mmdet"); _stub("mmdet.models", build_backbone=None)
    _stub("mmdet.models.backbones"); _stub("mmdet.models.backbones.resnet", BasicBlock=object)
    _stub("layers"); _stub("layers.backbones")
    _stub("layers.backbones.sam_encoder", build_sam_vit_b=None)
    # the op package: real python Function, stubbed native extension
    ops = _stub("ops"); ops.__path__ = [REF + "/ops"]
    vp_pkg = _stub("ops.voxel_pooling"); vp_pkg.__path__ = [REF + "/ops/voxel_pooling"]
    _stub("ops.voxel_pooling.voxel_pooling_ext", voxel_pooling_forward_wrapper=_naive_ext_forward)
    vp_pkg.voxel_pooling_ext = sys.modules["ops.voxel_pooling.voxel_pooling_ext"]
    spec = importlib.util.spec_from_file_location(
        "ops.voxel_pooling.voxel_pooling", REF + "/ops/voxel_pooling/voxel_pooling.py")
    vp = importlib.util.module_from_spec(spec); sys.modules[spec.name] = vp
    spec.loader.exec_module(vp)
    vp_pkg.voxel_pooling = vp.voxel_pooling
    spec = importlib.util.spec_from_file_location("ref_lss_fpn", REF + "/layers/backbones/lss_fpn.py")
    lss = importlib.util.module_from_spec(spec); spec.loader.exec_module(lss)
    return lss, vp


def load_dataset_helpers():
    import cv2
    tree = ast.parse(open(REF + "/dataset/nusc_mv_det_dataset.py").read())
    want = {"equation_plane", "get_denorm", "get_sensor2virtual", "get_reference_height"}
    ns = {"np": np, "math": math, "cv2": cv2}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            exec(compile(ast.Module([node], []), "ref_dataset", "exec"), ns)
    return ns


class _Shim:
    pass


def make_shim(lss, shape):
    o = _Shim()
    o.final_dim, o.downsample_factor, o.d_bound = shape.final_dim, shape.downsample, list(shape.d_bound)
    o.frustum = lss.LSSFPN.create_frustum(o)
    o.height2localtion = lambda *a: lss.LSSFPN.height2localtion(o, *a)
    rows = [shape.x_bound, shape.y_bound, shape.z_bound]       # lss_fpn.py:281-292 verbatim inputs
    o.voxel_size = torch.Tensor([r[2] for r in rows])
    o.voxel_coord = torch.Tensor([r[0] + r[2] / 2.0 for r in rows])
    o.voxel_num = torch.LongTensor([(r[1] - r[0]) / r[2] for r in rows])
    return o


def run_case(lss, vp, name, shape_name, batch, num_cams, seed, bda, with_pool, helpers=None):
    shape = get_shape(shape_name)
    shim = make_shim(lss, shape)
    mats = make_mats(shape, batch, num_cams, seed=seed, bda=bda)
    if helpers is not None:  # swap in sensor2virtual / reference height from the reference helpers
        for b in range(batch):
            for n in range(num_cams):
                e2s = np.linalg.inv(mats["sensor2ego"][b, n].numpy().astype(np.float64))
                dn = helpers["get_denorm"](e2s)
                mats["sensor2virtual"][b, n] = torch.from_numpy(helpers["get_sensor2virtual"](dn))
                mats["reference_heights"][b, n] = float(helpers["get_reference_height"](dn))
    geom = lss.LSSFPN.get_geometry(shim, mats["sensor2ego"], mats["sensor2virtual"], mats["intrin"],
                                   mats["ida"], mats["reference_heights"], mats["bda"])
    # lss_fpn.py:487-488, executed as written
    idx = ((geom - (shim.voxel_coord - shim.voxel_size / 2.0)) / shim.voxel_size).int()
    out = dict(
        shape=np.array(shape_name), batch=batch, num_cams=num_cams,
        frustum_u=shim.frustum[0, 0, :, 0].numpy(), frustum_v=shim.frustum[0, :, 0, 1].numpy(),
        frustum_z=shim.frustum[:, 0, 0, 2].numpy(),
        voxel_size=shim.voxel_size.numpy(), voxel_coord=shim.voxel_coord.numpy(),
        voxel_num=shim.voxel_num.numpy(),
        sensor2ego=mats["sensor2ego"].numpy(), sensor2virtual=mats["sensor2virtual"].numpy(),
        intrin=mats["intrin"].numpy(), ida=mats["ida"].numpy(),
        reference_heights=mats["reference_heights"].numpy(),
        bda=(mats["bda"].numpy() if mats["bda"] is not None else np.zeros(0, np.float32)),
        idx_sha256=np.array(hashlib.sha256(idx.numpy().tobytes()).hexdigest()),
        geom_sha256=np.array(hashlib.sha256(geom.contiguous().numpy().tobytes()).hexdigest()),
    )
    X, Y, Z = (int(v) for v in shim.voxel_num)
    i = idx.numpy()
    kept = (i[..., 0] >= 0) & (i[..., 0] < X) & (i[..., 1] >= 0) & (i[..., 1] < Y) & (i[..., 2] >= 0) & (i[..., 2] < Z)
    out["kept_count"] = int(kept.sum())
    if geom.numel() <= 200_000:
        out["geom"] = geom.contiguous().numpy()
        out["idx"] = idx.numpy()
    if with_pool:
        gen = torch.Generator().manual_seed(99 + seed)
        c = shape.channels
        bn = batch * num_cams
        height = torch.randn(bn, shape.D, shape.fH, shape.fW, generator=gen).softmax(1)
        ctx = torch.randn(bn, c, shape.fH, shape.fW, generator=gen)
        # lss_fpn.py:464-476,486 as written
        feat = height.unsqueeze(1) * ctx.unsqueeze(2)
        feat = feat.reshape(batch, num_cams, feat.shape[1], feat.shape[2], feat.shape[3], feat.shape[4])
        feat = feat.permute(0, 1, 3, 4, 5, 2).contiguous().requires_grad_(True)
        bev = vp.voxel_pooling(idx, feat, shim.voxel_num)            # reference autograd Function
        grad_bev = torch.randn(bev.shape, generator=gen)
        bev.contiguous().backward(grad_bev)
        out.update(height=height.numpy(), ctx=ctx.numpy(), bev=bev.detach().contiguous().numpy(),
                   grad_bev=grad_bev.numpy(), grad_feat=feat.grad.numpy())
    path = os.path.join(HERE, f"golden_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: points={geom.numel() // 3} kept={out['kept_count']} -> {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    torch.manual_seed(0)
    lss, vp = load_reference()
    helpers = load_dataset_helpers()
    run_case(lss, vp, "tiny_identity", "tiny", 2, 1, 11, "identity", True)
    run_case(lss, vp, "tiny_nobda_2cam", "tiny", 2, 2, 12, None, True)
    run_case(lss, vp, "small_random_bda", "small", 2, 1, 13, "random", True, helpers)
    run_case(lss, vp, "small_identity", "small", 1, 1, 14, "identity", False, helpers)
    # full-size cases: inputs + sha256 of the reference outputs only
    run_case(lss, vp, "dair_r50_hash", "dair_r50", 2, 1, 15, "identity", False)
    run_case(lss, vp, "rope3d_r50_hash", "rope3d_r50", 1, 1, 16, "identity", False, helpers)
    run_case(lss, vp, "sgv3d_bsm_r50_hash", "sgv3d_bsm_r50", 1, 1, 17, "identity", False)


if __name__ == "__main__":
    main()
