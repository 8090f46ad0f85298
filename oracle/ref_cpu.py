"""TEST INFRASTRUCTURE ONLY -- the reference's OWN Python for the lift-splat path, executed on the host CPU.

``make -C oracle ref`` byte-compiles ``layers/backbones/lss_fpn.py`` and ``ops/voxel_pooling/voxel_pooling.py``
UNMODIFIED from ``/root/reference`` into ``oracle/_ref/*.pyc`` (build outputs, git-ignored, shipped to the GPU box like
the compiled reference kernel).  This module loads that bytecode behind ``sys.modules`` stubs for the third-party
packages the reference imports but this image lacks (mmcv, mmdet, mmdet3d; same recipe as
tests/golden/make_golden.py) and exposes what ``bench.py``'s CPU arm times:

* ``LSSFPN.create_frustum`` / ``get_geometry`` / ``height2localtion``   (lss_fpn.py:325-401)  -- the reference's code
* the glue of ``_forward_single_sweep`` (lss_fpn.py:462-495): softmax, outer product, permute, quantise -- restated
  line by line below, because the method itself needs the backbone / mmcv modules
* ``voxel_pooling``: the reference op exists only as a CUDA kernel, so -- as BASELINE.json's north_star prescribes --
  its CPU stand-in is torch ``index_add_`` (oracle.lift_splat_oracle.voxel_pooling_forward)

Only tests/, __graft_entry__ and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PYC = os.path.join(_HERE, "_ref", "ref_lss_fpn.pycode")

_lss = None


def available() -> bool:
    return os.path.exists(PYC)


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """The reference's lss_fpn module (bytecode of the unmodified source)."""
    global _lss
    if _lss is not None:
        return _lss
    if not available():
        raise RuntimeError("oracle/_ref/ref_lss_fpn.pycode missing: run `make -C oracle ref` where /root/reference exists")
    _stub("mmcv"); _stub("mmcv.cnn", build_conv_layer=None)
    _stub("mmdet3d"); _stub("mmdet3d.models", build_neck=None)
    _stub("mmdet"); _stub("mmdet.models", build_backbone=None)
    _stub("mmdet.models.backbones"); _stub("mmdet.models.backbones.resnet", BasicBlock=object)
    _stub("layers"); _stub("layers.backbones")
    _stub("layers.backbones.sam_encoder", build_sam_vit_b=None)
    _stub("ops"); _stub("ops.voxel_pooling", voxel_pooling=None)   # CUDA-only op: replaced by index_add_ below
    loader = importlib.machinery.SourcelessFileLoader("ref_lss_fpn", PYC)
    spec = importlib.util.spec_from_loader("ref_lss_fpn", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _lss = mod
    return mod


class _Shim:
    """Carries the attributes the reference methods read from ``self`` (lss_fpn.py:281-294, 325-401)."""


def make_module(shape):
    lss = load()
    o = _Shim()
    o.final_dim, o.downsample_factor, o.d_bound = shape.final_dim, shape.downsample, list(shape.d_bound)
    o.frustum = lss.LSSFPN.create_frustum(o)
    o.height2localtion = lambda *a: lss.LSSFPN.height2localtion(o, *a)
    o.get_geometry = lambda *a: lss.LSSFPN.get_geometry(o, *a)
    rows = [shape.x_bound, shape.y_bound, shape.z_bound]
    o.voxel_size = torch.Tensor([r[2] for r in rows])
    o.voxel_coord = torch.Tensor([r[0] + r[2] / 2.0 for r in rows])
    o.voxel_num = torch.LongTensor([(r[1] - r[0]) / r[2] for r in rows])
    return o


def forward(module, height_logits, context, mats):
    """lss_fpn.py:462-495 on the CPU: the reference's get_geometry + the glue + index_add_ pooling."""
    from . import lift_splat_oracle as O
    b, nc = mats["sensor2ego"].shape[:2]
    height = height_logits.softmax(1)                                                  # :462
    feat = height.unsqueeze(1) * context.unsqueeze(2)                                  # :464-466
    feat = feat.reshape(b, nc, feat.shape[1], feat.shape[2], feat.shape[3], feat.shape[4])   # :469-476
    geom_xyz = module.get_geometry(mats["sensor2ego"], mats["sensor2virtual"], mats["intrin"], mats["ida"],
                                   mats["reference_heights"], mats.get("bda"))       # :478-485 (reference code)
    feat = feat.permute(0, 1, 3, 4, 5, 2)                                              # :486
    geom_xyz = ((geom_xyz - (module.voxel_coord - module.voxel_size / 2.0)) / module.voxel_size).int()   # :487-488
    bev, _ = O.voxel_pooling_forward(geom_xyz, feat.contiguous(), module.voxel_num)    # :490-491 (CUDA-only op)
    return bev.contiguous()                                                            # :494-495
