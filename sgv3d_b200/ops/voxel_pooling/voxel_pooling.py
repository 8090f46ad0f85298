"""Drop-in ``voxel_pooling(geom_xyz, input_features, voxel_num)`` operator.

Mirrors the reference operator ``ops/voxel_pooling/voxel_pooling.py:8-72`` (same name, argument
meaning, return layout and error behaviour) on top of the sm_100a kernels behind the C ABI
(``sgv3d_voxel_pooling_forward`` / ``_backward``).  Differences are internal only:

* deterministic: stable radix sort by voxel + ordered per-voxel sum instead of global float
  ``atomicAdd`` (ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:30-34);
* the gradient buffer is allocated in backward, not pre-zeroed in forward
  (voxel_pooling.py:29 materialises B*N*C floats during forward);
* backward is one gather kernel instead of a ``nonzero``/``index``/``index_put_`` chain
  (voxel_pooling.py:57-69, three host syncs).
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from ... import _native as N

__all__ = ["VoxelPooling", "voxel_pooling"]


def _voxel_num_to_ints(voxel_num):
    if isinstance(voxel_num, torch.Tensor):
        # the reference indexes the CUDA tensor element-wise (voxel_pooling.py:37-47): 5 host syncs.
        vals = voxel_num.tolist()
    else:
        vals = list(voxel_num)
    if len(vals) != 3:
        raise RuntimeError("voxel_num must hold 3 values (X, Y, Z)")
    return int(vals[0]), int(vals[1]), int(vals[2])


class VoxelPooling(Function):
    @staticmethod
    def forward(ctx, geom_xyz: torch.Tensor, input_features: torch.Tensor, voxel_num) -> torch.Tensor:
        """geom_xyz int32 [B, ..., 3] voxel coordinates, input_features fp32 [B, ..., C],
        voxel_num (X, Y, Z).  Returns the (B, C, Y, X) BEV feature map as a permuted view of a
        (B, Y, X, C) buffer, exactly like voxel_pooling.py:55."""
        assert geom_xyz.is_contiguous()                      # voxel_pooling.py:25
        assert input_features.is_contiguous()                # voxel_pooling.py:26
        if not (geom_xyz.is_cuda and input_features.is_cuda):
            raise RuntimeError("geom_xyz and input_features must be CUDA tensors")  # .cpp:12-18
        if geom_xyz.dtype != torch.int32:
            raise RuntimeError(f"expected geom_xyz of dtype int32, got {geom_xyz.dtype}")  # .cpp:30
        if input_features.dtype != torch.float32:
            raise RuntimeError(f"expected input_features of dtype float32, got {input_features.dtype}")
        ctx.mark_non_differentiable(geom_xyz)
        feat_shape = input_features.shape
        geom = geom_xyz.reshape(geom_xyz.shape[0], -1, geom_xyz.shape[-1])
        feat = input_features.reshape(geom.shape[0], -1, input_features.shape[-1])
        assert geom.shape[1] == feat.shape[1]                # voxel_pooling.py:33
        assert geom.shape[2] == 3
        b, n, c = feat.shape
        nx, ny, nz = _voxel_num_to_ints(voxel_num)
        out = feat.new_empty(b, ny, nx, c)
        pos_memo = geom.new_empty(b, n, 3)
        L = N.lib()
        ws_bytes = L.sgv3d_voxel_pooling_workspace_bytes(b, n, c, nx, ny, nz)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=feat.device)
        with torch.cuda.device(feat.device):
            N.check(L.sgv3d_voxel_pooling_forward(b, n, c, nx, ny, nz, N.ptr(geom), N.ptr(feat),
                                                  N.ptr(out), N.ptr(pos_memo), N.ptr(ws), ws_bytes,
                                                  N.current_stream()))
        ctx.save_for_backward(pos_memo)
        ctx.feat_shape = feat_shape
        ctx.grid = (nx, ny)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad_output_features):
        (pos_memo,) = ctx.saved_tensors
        b, n, _ = pos_memo.shape
        c = ctx.feat_shape[-1]
        nx, ny = ctx.grid
        g = grad_output_features
        if g.dtype != torch.float32:
            g = g.float()
        planar = g.is_contiguous()
        channels_last = g.stride(1) == 1 and g.permute(0, 2, 3, 1).is_contiguous()
        if not (planar or channels_last):
            g = g.contiguous()
            planar = True
        grad_feat = g.new_empty(b, n, c)
        L = N.lib()
        ws, ws_bytes = None, 0
        if not channels_last:
            ws_bytes = L.sgv3d_voxel_pooling_backward_workspace_bytes(b, c, nx, ny)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=g.device)
        sb, sc, sy, sx = g.stride()
        with torch.cuda.device(g.device):
            N.check(L.sgv3d_voxel_pooling_backward(b, n, c, nx, ny, N.ptr(g), sb, sc, sy, sx,
                                                   N.ptr(pos_memo), N.ptr(grad_feat), N.ptr(ws),
                                                   ws_bytes, N.current_stream()))
        return None, grad_feat.reshape(ctx.feat_shape), None     # voxel_pooling.py:69


voxel_pooling = VoxelPooling.apply
