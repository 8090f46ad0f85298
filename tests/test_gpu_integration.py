"""GPU: sgv3d_b200.integration.patch_view_transform on stand-ins for the reference's LSSFPN / BSMLSSFPN modules
(the real ones need mmcv / mmdet3d, absent here): same attribute and method names as layers/backbones/lss_fpn.py
and bsm_lss_fpn.py, tiny networks in front of the view transform.  The patched module must return what the
reference's _forward_single_sweep returns (lss_fpn.py:462-495, bsm_lss_fpn.py:523-559), computed here by the
oracle from the module's own intermediate tensors."""
import numpy as np
import pytest
import torch
from torch import nn

from oracle import c_oracle as CO
from oracle import lift_splat_oracle as O
from sgv3d_b200 import get_shape
from sgv3d_b200.integration import patch_view_transform, unpatch_view_transform
from sgv3d_b200.synthetic import make_mats
from sgv3d_b200.view_transform import build_frustum

pytestmark = pytest.mark.gpu


class _Base(nn.Module):
    """The registered buffers and constructor keys of lss_fpn.py:281-293."""

    def __init__(self, shape, stride, channels, is_train_height):
        super().__init__()
        rows = [shape.x_bound, shape.y_bound, shape.z_bound]
        self.register_buffer("voxel_size", torch.Tensor([r[2] for r in rows]))
        self.register_buffer("voxel_coord", torch.Tensor([r[0] + r[2] / 2.0 for r in rows]))
        self.register_buffer("voxel_num", torch.LongTensor([(r[1] - r[0]) / r[2] for r in rows]))
        self.register_buffer("frustum", build_frustum(shape.final_dim, stride, shape.d_bound))
        self.output_channels = channels
        self.height_channels = int(self.frustum.shape[0])
        self.is_train_height = is_train_height
        self.stride = stride

    def _forward_single_sweep(self, sweep_index, sweep_imgs, mats_dict):   # what patch_view_transform replaces
        raise NotImplementedError("the reference's own lift-splat (needs its CUDA extension)")

    def get_cam_feats(self, sweep_imgs):     # (B, sweeps, cams, 3, H, W) -> (B, sweeps, cams, F, fH, fW)
        b, s, n, c, h, w = sweep_imgs.shape
        f = self.stem(sweep_imgs.flatten(0, 2))
        return f.reshape(b, s, n, f.shape[1], f.shape[2], f.shape[3])


class LSSFPNStandIn(_Base):
    def __init__(self, shape, is_train_height=False):
        super().__init__(shape, shape.downsample, shape.channels, is_train_height)
        self.stem = nn.Conv2d(3, 6, self.stride, self.stride)
        self.assist_layer = nn.Conv2d(6, 4, 1)
        self.head = nn.Conv2d(6, self.height_channels + self.output_channels, 1)
        self.captured = None

    def _forward_height_net(self, feat, mats_dict):
        self.captured = self.head(feat)
        return self.captured


class BSMLSSFPNStandIn(_Base):
    def __init__(self, shape, is_train_height=False):
        super().__init__(shape, shape.downsample // 2, 80, is_train_height)   # bsm_lss_fpn.py:343
        self.stem = nn.Conv2d(3, 6, self.stride, self.stride)
        self.h_head, self.s_head, self.c_head = nn.Conv2d(6, self.height_channels, 1), nn.Conv2d(6, 7, 1), nn.Conv2d(6, 80, 1)
        self.captured = None

    def _forward_height_net(self, img_feats, mats_dict):
        f = img_feats[:, 0].flatten(0, 1)
        self.captured = (self.h_head(f), self.s_head(f) * 3.0, self.c_head(f), self.s_head(f))
        return self.captured


def _inputs(shape, batch):
    mats = make_mats(shape, batch, 1, seed=81, bda="identity")
    md = {"sensor2ego_mats": mats["sensor2ego"].unsqueeze(1).cuda(), "sensor2virtual_mats": mats["sensor2virtual"].unsqueeze(1).cuda(),
          "intrin_mats": mats["intrin"].unsqueeze(1).cuda(), "ida_mats": mats["ida"].unsqueeze(1).cuda(),
          "reference_heights": mats["reference_heights"].unsqueeze(1).cuda(), "bda_mat": mats["bda"].cuda()}
    g = torch.Generator().manual_seed(5)
    imgs = torch.randn(batch, 1, 1, 3, shape.final_dim[0], shape.final_dim[1], generator=g).cuda()
    return md, imgs


def _oracle_bev(mod, md, height_logits, context, grid):
    idx = mod._sgv3d_lift_splat.get_geometry_indices(md["sensor2ego_mats"][:, 0], md["sensor2virtual_mats"][:, 0],
                                                     md["intrin_mats"][:, 0], md["ida_mats"][:, 0],
                                                     md["reference_heights"][:, 0], md["bda_mat"]).cpu().numpy()
    return CO.lift_splat_forward64(idx, height_logits.detach().softmax(1).cpu().numpy(), context.detach().cpu().numpy(), *grid)


@pytest.mark.parametrize("train_height", [False, True])
def test_patched_lssfpn_returns_what_the_reference_call_site_returns(train_height):
    shape = get_shape("small")
    torch.manual_seed(3)
    mod = LSSFPNStandIn(shape, train_height).cuda()
    keys_before = set(mod.state_dict().keys())
    patch_view_transform(mod)
    assert set(mod.state_dict().keys()) == keys_before          # checkpoints keep loading
    md, imgs = _inputs(shape, 2)
    out = mod._forward_single_sweep(0, imgs, md)
    bev = out[0] if train_height else out
    if train_height:
        assert isinstance(out, tuple) and len(out[1]) == 2 and out[1][0].shape[1] == 4   # (assist, assist), lss_fpn.py:493
    d, c = mod.height_channels, mod.output_channels
    assert bev.shape == (2, c, shape.grid[1], shape.grid[0]) and bev.is_contiguous()
    want = _oracle_bev(mod, md, mod.captured[:, :d], mod.captured[:, d:d + c], shape.grid)
    np.testing.assert_allclose(bev.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    bev.sum().backward()                                           # gradients reach the networks in front
    assert mod.head.weight.grad is not None and torch.isfinite(mod.head.weight.grad).all()
    assert float(mod.head.weight.grad.abs().sum()) > 0
    unpatch_view_transform(mod)
    assert "_forward_single_sweep" not in mod.__dict__
    with pytest.raises(NotImplementedError):
        mod._forward_single_sweep(0, imgs, md)


@pytest.mark.parametrize("grad", [False, True])
def test_patched_bsm_module(grad):
    shape = get_shape("small")
    torch.manual_seed(4)
    mod = BSMLSSFPNStandIn(shape, is_train_height=True).cuda()
    patch_view_transform(mod)
    md, imgs = _inputs(shape, 2)
    with torch.set_grad_enabled(grad):
        bev, aux = mod._forward_single_sweep(0, imgs, md)
    assert aux[0] is mod.captured[3] and aux[1] is mod.captured[1]                      # (semantic0, semantic1), :558
    assert bev.shape == (2, 87, shape.grid[1], shape.grid[0]) and bev.is_contiguous()
    hl, sl, cx = (t.detach() for t in mod.captured[:3])
    feat = O.bsm_context(cx.cpu(), sl.cpu())
    want = _oracle_bev(mod, md, hl, feat, shape.grid)
    np.testing.assert_allclose(bev.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    if grad:
        bev.sum().backward()
        assert torch.isfinite(mod.c_head.weight.grad).all() and float(mod.c_head.weight.grad.abs().sum()) > 0


def test_patch_rejects_modules_that_are_not_lssfpn_like():
    with pytest.raises(RuntimeError, match="not an LSSFPN-like module"):
        patch_view_transform(nn.Linear(2, 2))
