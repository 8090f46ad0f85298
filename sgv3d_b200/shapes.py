"""Named lift-splat shapes lifted from the reference's experiment files.

Only the constants the view transform needs: image size, stride, height-bin bounds, channel
count and BEV grid.  Citations are relative to /root/reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple

__all__ = ["LiftSplatShape", "SHAPES", "get_shape"]


@dataclass(frozen=True)
class LiftSplatShape:
    name: str
    final_dim: Tuple[int, int]          # (H_in, W_in) of the network input image
    downsample: int                     # feature stride (LSSFPN: 16; BSMLSSFPN: 16 // 2 = 8)
    d_bound: Tuple[float, float, int]   # (z_lo, z_hi, D) height bins
    channels: int                       # C context channels
    x_bound: Tuple[float, float, float]
    y_bound: Tuple[float, float, float]
    z_bound: Tuple[float, float, float] = (-5.0, 3.0, 8.0)
    family: str = "dair"                # synthetic calibration family (see synthetic.py)
    source: str = field(default="", compare=False)

    @property
    def fH(self) -> int:
        return self.final_dim[0] // self.downsample

    @property
    def fW(self) -> int:
        return self.final_dim[1] // self.downsample

    @property
    def D(self) -> int:
        return int(self.d_bound[2])

    @property
    def grid(self) -> Tuple[int, int, int]:
        """(X, Y, Z) voxel counts, float->int truncation as torch.LongTensor does
        (layers/backbones/lss_fpn.py:289-292)."""
        return tuple(int((b[1] - b[0]) / b[2]) for b in (self.x_bound, self.y_bound, self.z_bound))

    @property
    def points_per_frame(self) -> int:
        return self.D * self.fH * self.fW

    # ---- algorithmic bytes per frame (SURVEY.md §8d / BASELINE.md §3) -----------------------
    def fused_forward_bytes(self, ctx_bytes: int = 4) -> int:
        X, Y, _ = self.grid
        return 4 * self.points_per_frame + ctx_bytes * self.channels * self.fH * self.fW \
            + 4 * self.channels * X * Y
    def fused_backward_bytes(self, ctx_bytes: int = 4, gctx_bytes: int = 4) -> int:
        X, Y, _ = self.grid
        px = self.fH * self.fW
        return 4 * self.channels * X * Y + 4 * self.points_per_frame + ctx_bytes * self.channels * px \
            + 4 * self.points_per_frame + gctx_bytes * self.channels * px
    def op_forward_bytes(self) -> int:
        X, Y, _ = self.grid
        n = self.points_per_frame
        return 4 * n * self.channels + 12 * n + 4 * self.channels * X * Y + 12 * n
    def op_backward_bytes(self) -> int:
        X, Y, _ = self.grid
        n = self.points_per_frame
        return 4 * self.channels * X * Y + 12 * n + 4 * n * self.channels


_G128 = dict(x_bound=(0.0, 102.4, 0.8), y_bound=(-51.2, 51.2, 0.8))
_G256 = dict(x_bound=(0.0, 102.4, 0.4), y_bound=(-51.2, 51.2, 0.4))
_G352 = dict(x_bound=(0.0, 140.8, 0.4), y_bound=(-70.4, 70.4, 0.4))

SHAPES = {s.name: s for s in [
    # north star: exps/bevheight/dair-v2x/bev_height_lss_r50_864_1536_128x128.py:38-48
    LiftSplatShape("dair_r50", (864, 1536), 16, (-2.0, 0.0, 90), 80, family="dair",
                   source="exps/bevheight/dair-v2x/bev_height_lss_r50_864_1536_128x128.py:38-48", **_G128),
    LiftSplatShape("dair_r50_256", (864, 1536), 16, (-2.0, 0.0, 90), 80, family="dair",
                   source="exps/bevheight/dair-v2x/bev_height_lss_r50_864_1536_256x256.py:34-43", **_G256),
    LiftSplatShape("rope3d_r50", (864, 1536), 16, (-2.0, 3.5, 90), 80, family="rope3d",
                   source="exps/bevheight/rope3d/bev_height_lss_r50_864_1536_128x128.py:44-53", **_G128),
    LiftSplatShape("rope3d_r101_256", (864, 1536), 16, (-2.0, 3.5, 180), 80, family="rope3d",
                   source="exps/bevheight/rope3d/bev_height_lss_r101_864_1536_256x256.py:45-54", **_G256),
    LiftSplatShape("rope3d_r101_140", (864, 1536), 16, (-0.5, 2.5, 90), 80, family="rope3d",
                   source="exps/bevheight/rope3d/bev_height_lss_r101_140.8_864_1536_256x256.py:45-54", **_G352),
    # SGV3D BSM: stride 8 (bsm_lss_fpn.py:343), C = 80 + 7 semantic (exps/sgv3d/...:40,88)
    LiftSplatShape("sgv3d_bsm_r50", (864, 1536), 8, (-2.0, 3.5, 90), 87, family="rope3d",
                   source="exps/sgv3d/bsm_bev_height_lss_r50_864_1536_128x128.py:42-73,88", **_G128),
    LiftSplatShape("sgv3d_bsm_r101", (864, 1536), 8, (-2.0, 3.5, 180), 87, family="rope3d",
                   source="exps/sgv3d/bsm_bev_height_lss_r101_864_1536_256x256.py:43-46", **_G256),
    # Rope3D native-resolution stress shape (derived)
    LiftSplatShape("rope3d_native", (1080, 1920), 16, (-2.0, 3.5, 90), 80, family="rope3d",
                   source="derived: final_dim=(1080,1920), floor division lss_fpn.py:329", **_G128),
    # tiny hand-checkable shape for golden fixtures / CPU tests
    LiftSplatShape("tiny", (48, 80), 16, (-2.0, 0.0, 4), 5, family="dair",
                   x_bound=(0.0, 102.4, 6.4), y_bound=(-51.2, 51.2, 6.4), source="test-only"),
    LiftSplatShape("small", (192, 320), 16, (-2.0, 3.5, 12), 7, family="rope3d",
                   x_bound=(0.0, 102.4, 1.6), y_bound=(-51.2, 51.2, 1.6), source="test-only"),
]}


def get_shape(name: str) -> LiftSplatShape:
    return SHAPES[name]
