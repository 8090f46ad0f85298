"""Host side of the fused lift-splat: plan / forward / backward over the C ABI, plus a module that
mirrors the view-transform half of the reference's ``LSSFPN`` / ``BSMLSSFPN``.

Reference call sites (relative to the reference repository root):
  layers/backbones/lss_fpn.py:462-495       LSSFPN._forward_single_sweep (after the height net)
  layers/backbones/bsm_lss_fpn.py:523-559   BSMLSSFPN._forward_single_sweep
  layers/backbones/lss_fpn.py:281-294       registered buffers (voxel_size/coord/num, frustum)
  layers/backbones/lss_fpn.py:325-401       create_frustum / height2localtion / get_geometry

The per-camera 4x4 prep (three inverses, two products: lss_fpn.py:361,367,392) runs in one library kernel that
restates torch's CUDA arithmetic bit for bit (verified against the torch calls at first use, which remain the
fallback), so the 16 floats per matrix that enter the per-point arithmetic are identical to the reference's
(SURVEY.md §7 hard part 1); everything per point / per pixel / per voxel happens in the sm_100a kernels.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch
from torch import nn
from torch.autograd import Function

from . import _native as N

__all__ = ["camera_matrices", "inverse4x4", "geometry_indices", "LiftSplatPlan", "lift_splat", "LiftSplat", "LiftSplatGraph",
           "build_frustum", "default_arith", "set_default_pipeline", "PIPELINE_AUTO", "PIPELINE_TILE", "PIPELINE_BLOCK"]

_DEFAULT_ARITH = N.ARITH_PAIR


def default_arith() -> int:
    """Evaluation order used for the per-point dot products unless overridden.
    PAIR (default) reproduces the reference executed on the GPU (torch CUDA ``matmul`` -> cuBLAS bmm on
    B200) bit for bit; SEQ reproduces the reference executed on CPU.  See DESIGN.md §3."""
    return _DEFAULT_ARITH


def set_default_arith(arith: int) -> None:
    global _DEFAULT_ARITH
    assert arith in (N.ARITH_SEQ, N.ARITH_FMA, N.ARITH_PAIR)
    _DEFAULT_ARITH = arith


PIPELINE_AUTO, PIPELINE_TILE, PIPELINE_BLOCK = 0, 1, 2
_DEFAULT_PIPELINE = PIPELINE_AUTO


def set_default_pipeline(pipeline: int) -> None:
    """Which kernels serve new plans: AUTO (LiftSplat._auto_pipeline: the pixel-block pipeline for small inference
    batches, the voxel-tile pipeline for training and large batches -- the B200 measurements in profiles/README.md),
    TILE, or BLOCK (the pixel-block pipeline: rows of up to 96 channels, D <= 255; error otherwise).  Plans built
    directly with ``LiftSplatPlan(...)`` and no ``pipeline=`` argument treat AUTO as TILE."""
    global _DEFAULT_PIPELINE
    assert pipeline in (PIPELINE_AUTO, PIPELINE_TILE, PIPELINE_BLOCK)
    _DEFAULT_PIPELINE = pipeline


def build_frustum(final_dim: Sequence[int], downsample_factor: int, d_bound: Sequence[float]) -> torch.Tensor:
    """(D, fH, fW, 4) fp32 buffer of (u, v, z_d, 1) -- same values as ``LSSFPN.create_frustum``
    (lss_fpn.py:325-348): torch fp32 ``linspace`` pixel centres and the float64 "DID" height bins
    z_d = d0 + (d/D)^1.5 (d1 - d0) cast to fp32."""
    in_h, in_w = final_dim
    f_h, f_w = in_h // downsample_factor, in_w // downsample_factor
    n_bins = int(d_bound[2])
    z = d_bound[0] + np.power(np.arange(n_bins) / n_bins, 1.5) * (d_bound[1] - d_bound[0])
    z = torch.tensor(z, dtype=torch.float).view(n_bins, 1, 1).expand(n_bins, f_h, f_w)
    u = torch.linspace(0, in_w - 1, f_w, dtype=torch.float).view(1, 1, f_w).expand(n_bins, f_h, f_w)
    v = torch.linspace(0, in_h - 1, f_h, dtype=torch.float).view(1, f_h, 1).expand(n_bins, f_h, f_w)
    return torch.stack((u, v, z, torch.ones_like(z)), -1)


def _inverse(x: torch.Tensor) -> torch.Tensor:
    """``torch.inverse`` without its host synchronisation: ``linalg.inv_ex`` is the routine
    ``torch.inverse`` / ``Tensor.inverse`` dispatch to, minus the ``info`` check that forces a
    device->host copy (bit-identical results; tests/test_gpu_lift_splat.py checks that)."""
    return torch.linalg.inv_ex(x, check_errors=False).inverse


_INVERSE_KERNEL_OK: Dict[int, bool] = {}


def inverse4x4(*mats: torch.Tensor):
    """Inverses of up to three equally shaped (..., 4, 4) fp32 CUDA tensors in ONE launch
    (``sgv3d_inverse4x4``: the rounding sequence of torch's CUDA ``inverse``, restated)."""
    assert 1 <= len(mats) <= 3
    a = [m.contiguous() for m in mats]
    n = a[0].numel() // 16
    out = torch.empty((len(a),) + tuple(a[0].shape), dtype=torch.float32, device=a[0].device)
    ptrs = [N.ptr(t) for t in a] + [0] * (3 - len(a))
    outs = [N.ptr(out[i]) for i in range(len(a))] + [0] * (3 - len(a))
    with torch.cuda.device(a[0].device):
        N.check(N.lib().sgv3d_inverse4x4(n, *ptrs, *outs, N.current_stream()))
    return tuple(out[i] for i in range(len(a)))


def _inverse_kernel_verified(device) -> bool:
    """One-time check per device that ``sgv3d_inverse4x4`` reproduces this installation's
    ``torch.linalg.inv_ex`` bit for bit (it restates cuBLAS' batched LU / solve arithmetic, which a different
    library version could change); on any difference the torch routine stays in use."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    ok = _INVERSE_KERNEL_OK.get(idx)
    if ok is None:
        if torch.cuda.is_current_stream_capturing():
            return False          # decided by the first eager call
        g = torch.Generator().manual_seed(20260607)
        a = torch.randn(384, 4, 4, generator=g)
        a[:128] += 4.0 * torch.eye(4)
        a[128:256] *= torch.tensor([1e-2, 1.0, 30.0, 2e3]).view(1, 1, 4)
        a[256:320, 3] = torch.tensor([0.0, 0.0, 0.0, 1.0])        # affine, like every calibration matrix
        a = a.to(device)
        want = _inverse(a)
        (got,) = inverse4x4(a)
        ok = bool(torch.equal(want.view(torch.int32), got.view(torch.int32)))
        _INVERSE_KERNEL_OK[idx] = ok
        if not ok:
            import warnings
            warnings.warn("sgv3d_inverse4x4 differs from torch.linalg.inv_ex on this installation; "
                          "using torch's routine for the per-camera inverses")
    return ok


_CAMERA_PREP_OK: Dict[tuple, bool] = {}


def _camera_prep(sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat):
    """``sgv3d_camera_prep``: the three inverses and the two products of lss_fpn.py:361,367,392 in one launch."""
    shape = tuple(ida_mat.shape)
    n = ida_mat.numel() // 16
    out = torch.empty((3,) + shape, dtype=torch.float32, device=ida_mat.device)
    a = [t.contiguous() for t in (ida_mat, intrin_mat, sensor2virtual_mat, sensor2ego_mat)]
    with torch.cuda.device(ida_mat.device):
        N.check(N.lib().sgv3d_camera_prep(n, N.ARITH_SEQ if n == 1 else N.ARITH_FMA, *[N.ptr(t) for t in a],
                                          N.ptr(out[0]), N.ptr(out[1]), N.ptr(out[2]), N.current_stream()))
    return out[0], out[1], out[2]


def _camera_prep_verified(device, shape) -> bool:
    """One-time check per device and batch shape that the fused prep kernel reproduces torch's ``inverse`` and
    batched ``matmul`` bit for bit (torch picks its matmul kernel, and with it the rounding order, by batch count:
    tools/probe_matmul.py); on any difference the torch calls stay in use."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, tuple(shape))
    ok = _CAMERA_PREP_OK.get(key)
    if ok is None:
        if torch.cuda.is_current_stream_capturing():
            return False          # decided by the first eager call
        if not _inverse_kernel_verified(device):
            _CAMERA_PREP_OK[key] = False
            return False
        g = torch.Generator().manual_seed(20260608)
        ok = True
        for rep in range(4):     # small batches hold few matrices: several draws
            mats = []
            for k in range(4):
                a = torch.randn(*shape, generator=g) + 3.0 * torch.eye(4)
                a[..., 3, :] = torch.tensor([0.0, 0.0, 0.0, 1.0])
                mats.append(a.to(device))
            ida, intrin, s2v, s2e = mats
            want = (_inverse(ida), s2v.matmul(_inverse(intrin)), s2e.matmul(_inverse(s2v)))
            got = _camera_prep(s2e, s2v, intrin, ida)
            ok = ok and all(torch.equal(w.view(torch.int32), g_.view(torch.int32)) for w, g_ in zip(want, got))
        ok = bool(ok)
        _CAMERA_PREP_OK[key] = ok
    return ok


def camera_matrices(sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat):
    """``ida.inverse()``, ``sensor2virtual @ inverse(intrin)``, ``sensor2ego @ inverse(sensor2virtual)``
    evaluated with the reference's own torch routines (lss_fpn.py:392,361,367), shapes (B, Nc, 4, 4)."""
    same = ida_mat.shape == intrin_mat.shape == sensor2virtual_mat.shape and ida_mat.dim() >= 3
    f32 = ida_mat.dtype == intrin_mat.dtype == sensor2virtual_mat.dtype == sensor2ego_mat.dtype == torch.float32
    if ida_mat.is_cuda and same and f32 and sensor2ego_mat.shape == ida_mat.shape and ida_mat.numel() > 0 and \
            _camera_prep_verified(ida_mat.device, ida_mat.shape):
        # inverses AND products in one launch, bit-identical to the torch calls of the reference (verified above
        # and by tests/test_gpu_lift_splat.py::test_camera_prep_kernel_is_bit_identical_to_torch)
        return _camera_prep(sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat)
    if ida_mat.is_cuda and same and ida_mat.numel() > 0 and \
            ida_mat.dtype == intrin_mat.dtype == sensor2virtual_mat.dtype == torch.float32 and \
            _inverse_kernel_verified(ida_mat.device):
        # the three batched LU inverses of the reference (lss_fpn.py:392,361,367) in one launch instead of torch's
        # cat + getrf + laswp + 2 x trsm + identity / pivot set-up kernels; bit-identical (verified above and by
        # tests/test_gpu_lift_splat.py::test_inverse4x4_kernel_is_bit_identical_to_torch)
        ida_inv, intrin_inv, s2v_inv = inverse4x4(ida_mat, intrin_mat, sensor2virtual_mat)
    elif ida_mat.is_cuda and same:
        # one batched LU for the three inverses instead of three: the batched routine treats every 4x4
        # independently, so each inverse keeps the bits of its own call (tools/probe_prep.py;
        # tests/test_gpu_lift_splat.py::test_stacked_inverse_is_bit_identical) at a third of the launches.
        # The two products stay separate calls: cuBLAS picks its kernel by batch count.
        b = ida_mat.shape[0]
        inv = _inverse(torch.cat((ida_mat, intrin_mat, sensor2virtual_mat), 0))
        ida_inv, intrin_inv, s2v_inv = inv[:b], inv[b:2 * b], inv[2 * b:]
    else:
        ida_inv, intrin_inv, s2v_inv = _inverse(ida_mat), _inverse(intrin_mat), _inverse(sensor2virtual_mat)
    m_virtual = sensor2virtual_mat.matmul(intrin_inv)
    m_ego = sensor2ego_mat.matmul(s2v_inv)
    return ida_inv, m_virtual, m_ego


def _f32c(t: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
    if t is None:
        return None
    return t.to(device=device, dtype=torch.float32).contiguous()


class _GridConst:
    """Host-side constants that never change for a module: frustum axes on the device and the
    quantisation lower bound / voxel size as host floats (reading them from CUDA buffers every call
    would synchronise, as lss_fpn.py:487-491 does)."""

    def __init__(self, frustum, voxel_coord, voxel_size, device):
        fr = frustum.to(device)
        self.D, self.fH, self.fW = (int(s) for s in fr.shape[:3])
        # the three axes of the frustum buffer; looked up, never recomputed
        self.u = fr[0, 0, :, 0].contiguous()
        self.v = fr[0, :, 0, 1].contiguous()
        self.z = fr[:, 0, 0, 2].contiguous()
        # lss_fpn.py:487-488: lower = voxel_coord - voxel_size / 2.0 in fp32 tensor arithmetic
        vc, vs = voxel_coord.detach().float().cpu(), voxel_size.detach().float().cpu()
        self.lower = N.host_f32x3((vc - vs / 2.0).tolist())
        self.size = N.host_f32x3(vs.tolist())
        self.device = device


class _Geometry:
    """Device-resident operands of the per-point geometry for one batch of cameras."""

    def __init__(self, frustum, sensor2ego, sensor2virtual, intrin, ida, reference_heights, bda,
                 voxel_coord, voxel_size, grid_const: Optional[_GridConst] = None):
        dev = sensor2ego.device
        if dev.type != "cuda":
            raise RuntimeError("sgv3d_b200 runs on CUDA tensors only (no CPU fallback)")
        self.device = dev
        self.B, self.Nc = int(sensor2ego.shape[0]), int(sensor2ego.shape[1])
        gc = grid_const if grid_const is not None and grid_const.device == dev else \
            _GridConst(frustum, voxel_coord, voxel_size, dev)
        self.gc = gc
        self.D, self.fH, self.fW = gc.D, gc.fH, gc.fW
        ida_inv, m_virtual, m_ego = camera_matrices(sensor2ego, sensor2virtual, intrin, ida)
        self.ida_inv, self.m_virtual, self.m_ego = (_f32c(t, dev) for t in (ida_inv, m_virtual, m_ego))
        # private copies: update_calibration() overwrites these in place and must never write through to the
        # caller's reference_heights / bda_mat tensors
        self.ref_h = _f32c(reference_heights, dev).reshape(-1).clone()
        self.bda = _f32c(bda, dev).clone() if bda is not None else None

    def pointer_args(self):
        gc = self.gc
        return [N.ptr(gc.u), N.ptr(gc.v), N.ptr(gc.z), N.ptr(self.ida_inv), N.ptr(self.m_virtual),
                N.ptr(self.m_ego), N.ptr(self.bda), N.ptr(self.ref_h), gc.lower, gc.size]


def geometry_indices(frustum, sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat, reference_heights,
                     bda_mat, voxel_coord, voxel_size, arith: Optional[int] = None, return_xyz: bool = False,
                     grid_const: "Optional[_GridConst]" = None):
    """int32 (B, Nc, D, fH, fW, 3) voxel indices == ``((get_geometry(...) - lower) / size).int()``
    (lss_fpn.py:478-488), computed by one kernel; optionally also the fp32 ego coordinates."""
    g = _Geometry(frustum, sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat, reference_heights,
                  bda_mat, voxel_coord, voxel_size, grid_const)
    idx = torch.empty(g.B, g.Nc, g.D, g.fH, g.fW, 3, dtype=torch.int32, device=g.device)
    xyz = torch.empty(idx.shape, dtype=torch.float32, device=g.device) if return_xyz else None
    with torch.cuda.device(g.device):
        N.check(N.lib().sgv3d_geometry_quantize(
            default_arith() if arith is None else arith, g.B, g.Nc, g.D, g.fH, g.fW, *g.pointer_args(),
            N.ptr(idx), N.ptr(xyz), N.current_stream()))
    return (idx, xyz) if return_xyz else idx


def _camera_block_stride(t: torch.Tensor, channels: int, fh: int, fw: int, what: str) -> int:
    """Element stride between consecutive cameras of a (BN, channels, fH, fW) tensor whose per-camera
    block is dense (a contiguous tensor, or a channel slice of the height net's output)."""
    assert tuple(t.shape[1:]) == (channels, fh, fw), (what, tuple(t.shape), (channels, fh, fw))
    if t.shape[0] > 0 and t.stride()[1:] != (fh * fw, fw, 1):
        raise RuntimeError(f"{what}: per-camera [C][fH][fW] block must be dense, got strides {t.stride()}")
    return int(t.stride(0)) if t.shape[0] > 1 else channels * fh * fw


class LiftSplatPlan:
    """Sorted voxel-run index for one batch of calibrations (``sgv3d_lift_splat_plan``).

    Depends only on the matrices and the grid, so a static roadside camera can build it once and
    reuse it for every frame; in training it is rebuilt per step."""

    def __init__(self, frustum, sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat, reference_heights,
                 bda_mat, voxel_coord, voxel_size, voxel_num: Sequence[int], channels: int,
                 ctx_dtype: torch.dtype = torch.float32, arith: Optional[int] = None,
                 grid_const: Optional[_GridConst] = None, pipeline: Optional[int] = None,
                 channels_last: bool = False):
        g = _Geometry(frustum, sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat, reference_heights,
                      bda_mat, voxel_coord, voxel_size, grid_const)
        self.geometry = g
        self.device = g.device
        if ctx_dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError(f"context dtype {ctx_dtype} unsupported (float32 or bfloat16)")
        self.ctx_dtype = ctx_dtype
        nx, ny, nz = (int(v) for v in voxel_num)
        self.desc = N.LiftSplatDesc(B=g.B, Nc=g.Nc, D=g.D, fH=g.fH, fW=g.fW, C=int(channels), X=nx, Y=ny, Z=nz,
                                    arith=default_arith() if arith is None else arith,
                                    ctx_dtype=N.DTYPE_BF16 if ctx_dtype == torch.bfloat16 else N.DTYPE_F32)
        self.desc.reserved[0] = _DEFAULT_PIPELINE if pipeline is None else pipeline
        # BEV map (and the gradient the backward reads) in torch.channels_last memory order -- same logical
        # (B, C, Y, X) tensor, no transposed copy on either side (include/sgv3d_b200.h: reserved[1] = 2)
        self.channels_last = bool(channels_last)
        if self.channels_last and (int(channels) % 16 != 0 or int(channels) > 96 or self.desc.reserved[0] == PIPELINE_BLOCK):
            raise RuntimeError("channels_last BEV maps need the voxel-tile pipeline and 16, 32, ..., 96 channels")
        L = N.lib()
        self.ws_bytes = L.sgv3d_lift_splat_workspace_bytes(self.desc)
        if g.B > 0 and self.ws_bytes == 0:
            N.check(1)
        self.ws = torch.empty(max(self.ws_bytes, 1), dtype=torch.uint8, device=g.device)
        self.rebuild()

    def update_calibration(self, sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat, reference_heights,
                           bda_mat) -> None:
        """New matrices for the same batch shape: the derived per-camera operands are overwritten in place
        and the plan is rebuilt into the same workspace."""
        g = self.geometry
        ida_inv, m_virtual, m_ego = camera_matrices(sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat)
        g.ida_inv.copy_(ida_inv.reshape(g.ida_inv.shape))
        g.m_virtual.copy_(m_virtual.reshape(g.m_virtual.shape))
        g.m_ego.copy_(m_ego.reshape(g.m_ego.shape))
        g.ref_h.copy_(reference_heights.reshape(-1))
        if (g.bda is None) != (bda_mat is None):
            raise RuntimeError("update_calibration cannot add or remove the BDA matrix")
        if bda_mat is not None:
            g.bda.copy_(bda_mat)
        self.rebuild()

    def rebuild(self) -> None:
        g = self.geometry
        with torch.cuda.device(self.device):
            N.check(N.lib().sgv3d_lift_splat_plan(self.desc, *g.pointer_args(), N.ptr(self.ws), self.ws_bytes,
                                                  N.current_stream()))

    def _call_desc(self, height, context, logits, g_height=None, g_ctx=None) -> "N.LiftSplatDesc":
        d = self.desc
        self._check_inputs(height, context)
        c = N.LiftSplatDesc.from_buffer_copy(d)
        c.height_is_logits = 1 if logits else 0
        c.reserved[1] = 2 if self.channels_last else 0
        c.height_batch_stride = _camera_block_stride(height, d.D, d.fH, d.fW, "height")
        c.ctx_batch_stride = _camera_block_stride(context, d.C, d.fH, d.fW, "context")
        if g_height is not None:
            c.grad_height_batch_stride = _camera_block_stride(g_height, d.D, d.fH, d.fW, "grad_height")
            c.grad_ctx_batch_stride = _camera_block_stride(g_ctx, d.C, d.fH, d.fW, "grad_context")
        return c

    # -- raw entry points (no autograd) -------------------------------------------------------------
    def forward(self, height: torch.Tensor, context: torch.Tensor, logits: bool = False) -> torch.Tensor:
        """``height``: probabilities (or raw logits with ``logits=True``: the softmax over D is fused)."""
        d = self._call_desc(height, context, logits)
        bev = torch.empty(d.B, d.C, d.Y, d.X, dtype=torch.float32, device=self.device,
                          memory_format=torch.channels_last if self.channels_last else torch.contiguous_format)
        with torch.cuda.device(self.device):
            N.check(N.lib().sgv3d_lift_splat_forward(d, N.ptr(height), N.ptr(context), N.ptr(bev), N.ptr(self.ws),
                                                     self.ws_bytes, N.current_stream()))
        return bev

    def forward_bsm(self, height: torch.Tensor, context: torch.Tensor, semantic_logits: torch.Tensor,
                    threshold: float = 0.45, logits: bool = True) -> torch.Tensor:
        """BSMLSSFPN forward with the context assembly of bsm_lss_fpn.py:524-529 fused (inference, no autograd):
        ``context`` (BN, C - Cs, fH, fW) and ``semantic_logits`` (BN, Cs, fH, fW) are consumed in place; the
        masked (BN, C, fH, fW) tensor is never built."""
        d = self.desc
        bn, cs = d.B * d.Nc, int(semantic_logits.shape[1])
        cc = d.C - cs
        if self.ctx_dtype != torch.float32:
            raise RuntimeError("forward_bsm: float32 context only")
        if self.channels_last:
            raise RuntimeError("forward_bsm: contiguous (B, C, Y, X) BEV maps only")
        for name, t, ch in (("height", height, d.D), ("context", context, cc), ("semantic_logits", semantic_logits, cs)):
            if not t.is_cuda or t.dtype != torch.float32:
                raise RuntimeError(f"forward_bsm: {name} must be a float32 CUDA tensor")
            assert tuple(t.shape) == (bn, ch, d.fH, d.fW), (name, tuple(t.shape), (bn, ch, d.fH, d.fW))
        height, context, semantic_logits = _dense_blocks(height), _dense_blocks(context), _dense_blocks(semantic_logits)
        c = N.LiftSplatDesc.from_buffer_copy(d)
        c.height_is_logits = 1 if logits else 0
        c.height_batch_stride = _camera_block_stride(height, d.D, d.fH, d.fW, "height")
        c.ctx_batch_stride = _camera_block_stride(context, cc, d.fH, d.fW, "context")
        sem_stride = _camera_block_stride(semantic_logits, cs, d.fH, d.fW, "semantic_logits")
        bev = torch.empty(d.B, d.C, d.Y, d.X, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(N.lib().sgv3d_lift_splat_forward_bsm(c, N.ptr(height), N.ptr(context), N.ptr(semantic_logits), cs,
                                                         sem_stride, float(threshold), N.ptr(bev), N.ptr(self.ws),
                                                         self.ws_bytes, N.current_stream()))
        return bev

    def uses_block_pipeline(self) -> bool:
        return bool(N.lib().sgv3d_lift_splat_uses_block_pipeline(self.desc))

    def backward_bsm(self, grad_bev: torch.Tensor, height: torch.Tensor, context: torch.Tensor,
                     semantic_logits: torch.Tensor, threshold: float = 0.45, logits: bool = True):
        """Gradients of the BSMLSSFPN call site with the context assembly fused (bsm_lss_fpn.py:523-541 under autograd):
        (grad_height [logits], grad_context (BN, C - Cs, fH, fW), grad_semantic_logits (BN, Cs, fH, fW))."""
        d = self.desc
        bn, cs = d.B * d.Nc, int(semantic_logits.shape[1])
        cc = d.C - cs
        g = grad_bev.float().contiguous()
        assert tuple(g.shape) == (d.B, d.C, d.Y, d.X)
        height, context, semantic_logits = _dense_blocks(height), _dense_blocks(context), _dense_blocks(semantic_logits)
        for name, t, ch in (("height", height, d.D), ("context", context, cc), ("semantic_logits", semantic_logits, cs)):
            if not t.is_cuda or t.dtype != torch.float32:
                raise RuntimeError(f"backward_bsm: {name} must be a float32 CUDA tensor")
            assert tuple(t.shape) == (bn, ch, d.fH, d.fW), (name, tuple(t.shape), (bn, ch, d.fH, d.fW))
        g_height = torch.empty(bn, d.D, d.fH, d.fW, dtype=torch.float32, device=self.device)
        g_ctx = torch.empty(bn, cc, d.fH, d.fW, dtype=torch.float32, device=self.device)
        g_sem = torch.empty(bn, cs, d.fH, d.fW, dtype=torch.float32, device=self.device)
        c = N.LiftSplatDesc.from_buffer_copy(d)
        c.height_is_logits = 1 if logits else 0
        c.height_batch_stride = _camera_block_stride(height, d.D, d.fH, d.fW, "height")
        c.ctx_batch_stride = _camera_block_stride(context, cc, d.fH, d.fW, "context")
        sem_stride = _camera_block_stride(semantic_logits, cs, d.fH, d.fW, "semantic_logits")
        with torch.cuda.device(self.device):
            N.check(N.lib().sgv3d_lift_splat_backward_bsm(c, N.ptr(g), N.ptr(height), N.ptr(context),
                                                          N.ptr(semantic_logits), cs, sem_stride, float(threshold),
                                                          N.ptr(g_height), N.ptr(g_ctx), N.ptr(g_sem), N.ptr(self.ws),
                                                          self.ws_bytes, N.current_stream()))
        return g_height, g_ctx, g_sem

    def backward(self, grad_bev: torch.Tensor, height: torch.Tensor, context: torch.Tensor, logits: bool = False,
                 out_height: Optional[torch.Tensor] = None, out_context: Optional[torch.Tensor] = None):
        """Gradients w.r.t. ``height`` (w.r.t. the logits when ``logits=True``) and ``context``; optionally
        written straight into caller-provided (possibly strided) buffers."""
        # (a channels_last plan reads the gradient in channels_last order: no copy when the BEV trunk's backward
        # produced it that way, which it does when it consumed the channels_last map)
        g = grad_bev.float().contiguous(memory_format=torch.channels_last if self.channels_last
                                        else torch.contiguous_format)
        d0 = self.desc
        assert tuple(g.shape) == (d0.B, d0.C, d0.Y, d0.X)
        bn = d0.B * d0.Nc
        g_height = out_height if out_height is not None else \
            torch.empty(bn, d0.D, d0.fH, d0.fW, dtype=torch.float32, device=self.device)
        g_ctx = out_context if out_context is not None else \
            torch.empty(bn, d0.C, d0.fH, d0.fW, dtype=torch.float32, device=self.device)
        assert g_height.dtype == torch.float32 and g_ctx.dtype == torch.float32
        d = self._call_desc(height, context, logits, g_height, g_ctx)
        with torch.cuda.device(self.device):
            N.check(N.lib().sgv3d_lift_splat_backward(d, N.ptr(g), N.ptr(height), N.ptr(context), N.ptr(g_height),
                                                      N.ptr(g_ctx), N.ptr(self.ws), self.ws_bytes,
                                                      N.current_stream()))
        return g_height, g_ctx

    def expand(self) -> torch.Tensor:
        """int32 (B, Nc, D, fH, fW): voxel id y*X+x per point, -1 for dropped points (parity/debug)."""
        d = self.desc
        vox = torch.empty(d.B, d.Nc, d.D, d.fH, d.fW, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(N.lib().sgv3d_lift_splat_plan_expand(d, N.ptr(vox), N.ptr(self.ws), self.ws_bytes,
                                                         N.current_stream()))
        return vox

    def _check_inputs(self, height, context):
        d = self.desc
        bn = d.B * d.Nc
        if not (height.is_cuda and context.is_cuda):
            raise RuntimeError("height and context must be CUDA tensors")
        if height.dtype != torch.float32:
            raise RuntimeError(f"expected height of dtype float32, got {height.dtype}")
        if context.dtype != self.ctx_dtype:
            raise RuntimeError(f"expected context of dtype {self.ctx_dtype}, got {context.dtype}")
        assert tuple(height.shape) == (bn, d.D, d.fH, d.fW), (tuple(height.shape), (bn, d.D, d.fH, d.fW))
        assert tuple(context.shape) == (bn, d.C, d.fH, d.fW), (tuple(context.shape), (bn, d.C, d.fH, d.fW))


def _dense_blocks(t: torch.Tensor) -> torch.Tensor:
    """Return ``t`` itself when every camera's [C][fH][fW] block is dense (e.g. a channel slice of the
    height net's output), else a contiguous copy."""
    return t if t.stride()[1:] == (t.shape[2] * t.shape[3], t.shape[3], 1) else t.contiguous()


class _LiftSplatFunction(Function):
    @staticmethod
    def forward(ctx, height, context, plan: LiftSplatPlan, logits: bool):
        height, context = _dense_blocks(height), _dense_blocks(context)
        ctx.plan, ctx.logits = plan, logits
        ctx.save_for_backward(height, context)
        return plan.forward(height, context, logits)

    @staticmethod
    def backward(ctx, grad_bev):
        height, context = ctx.saved_tensors
        g_height, g_ctx = ctx.plan.backward(grad_bev, height, context, ctx.logits)
        return g_height, g_ctx.to(context.dtype), None, None


class _LiftSplatHeadFunction(Function):
    """Whole LSSFPN call site on the height net's output tensor (BN, D + C, fH, fW): height logits and
    context are consumed in place (no slicing copies) and the gradient is written straight into one
    (BN, D + C, fH, fW) buffer (no ``cat``)."""

    @staticmethod
    def forward(ctx, height_feature, plan: LiftSplatPlan, d: int, c: int):
        hf = height_feature if height_feature.is_contiguous() else height_feature.contiguous()
        ctx.plan, ctx.dc = plan, (d, c)
        ctx.save_for_backward(hf)
        return plan.forward(hf[:, :d], hf[:, d:d + c], logits=True)

    @staticmethod
    def backward(ctx, grad_bev):
        (hf,) = ctx.saved_tensors
        d, c = ctx.dc
        g = torch.empty_like(hf) if hf.shape[1] == d + c else torch.zeros_like(hf)
        ctx.plan.backward(grad_bev, hf[:, :d], hf[:, d:d + c], logits=True,
                          out_height=g[:, :d], out_context=g[:, d:d + c])
        return g, None, None, None


class _LiftSplatBsmFunction(Function):
    """BSMLSSFPN call site under autograd with the context assembly (7-way semantic softmax, concat, background mask:
    bsm_lss_fpn.py:524-529) fused into both passes: the 87-channel masked tensor is never built, masked pixels are
    skipped, and the semantic-softmax backward runs inside the backward kernel."""

    @staticmethod
    def forward(ctx, height_logits, semantic_logits, context, plan: LiftSplatPlan, threshold: float):
        h, s, c = _dense_blocks(height_logits), _dense_blocks(semantic_logits), _dense_blocks(context)
        ctx.plan, ctx.threshold = plan, threshold
        ctx.save_for_backward(h, s, c)
        return plan.forward_bsm(h, c, s, threshold, logits=True)

    @staticmethod
    def backward(ctx, grad_bev):
        h, s, c = ctx.saved_tensors
        g_h, g_c, g_s = ctx.plan.backward_bsm(grad_bev, h, c, s, ctx.threshold, logits=True)
        return g_h, g_s, g_c, None, None


def lift_splat(height: torch.Tensor, context: torch.Tensor, plan: LiftSplatPlan, logits: bool = False) -> torch.Tensor:
    """(B, C, Y, X) fp32 contiguous BEV map: ``voxel_pooling(idx, (height (x) context) permuted)``
    of lss_fpn.py:464-495 without the frustum tensor.  ``height`` holds the softmax-ed height-bin
    probabilities, or the raw logits with ``logits=True`` (softmax of lss_fpn.py:462 fused).
    Differentiable w.r.t. height and context."""
    return _LiftSplatFunction.apply(height, context, plan, logits)


class LiftSplat(nn.Module):
    """View-transform half of ``LSSFPN`` / ``BSMLSSFPN``: same constructor keys, same registered
    buffers (so ``state_dict`` entries line up with lss_fpn.py:281-293), same per-sweep maths.

    ``forward_single_sweep`` consumes what the height net produced and returns what
    ``_forward_single_sweep`` returns for the BEV map (lss_fpn.py:494-495)."""

    def __init__(self, x_bound, y_bound, z_bound, d_bound, final_dim, downsample_factor, output_channels,
                 is_bsm: bool = False, arith: Optional[int] = None, cache_plan: bool = False,
                 bev_channels_last: bool = False):
        super().__init__()
        # True: the BEV map is returned with torch.channels_last strides (same values, same shape) for a BEV trunk
        # that runs in that memory format, and its gradient is consumed in that order -- the forward's transpose
        # through shared memory and the backward's gradient-row copy disappear (LSSFPN call site, 16 .. 96 channels)
        self.bev_channels_last = bool(bev_channels_last) and not is_bsm
        # BSMLSSFPN halves the stride of the lifted feature map (bsm_lss_fpn.py:343)
        self.downsample_factor = downsample_factor // 2 if is_bsm else downsample_factor
        self.is_bsm = is_bsm
        self.d_bound = list(d_bound)
        self.final_dim = tuple(final_dim)
        self.output_channels = output_channels
        self.arith = arith
        self.cache_plan = cache_plan
        rows = [x_bound, y_bound, z_bound]
        self.register_buffer("voxel_size", torch.Tensor([r[2] for r in rows]))
        self.register_buffer("voxel_coord", torch.Tensor([r[0] + r[2] / 2.0 for r in rows]))
        self.register_buffer("voxel_num", torch.LongTensor([(r[1] - r[0]) / r[2] for r in rows]))
        self.register_buffer("frustum", build_frustum(self.final_dim, self.downsample_factor, self.d_bound))
        self.height_channels = int(self.frustum.shape[0])
        # host copies, so that no call has to read a CUDA scalar (lss_fpn.py:491 syncs every call)
        self._grid = tuple(int(v) for v in self.voxel_num.tolist())
        self._plan_cache: Dict[tuple, LiftSplatPlan] = {}
        self._grid_const: Optional[_GridConst] = None

    @classmethod
    def from_buffers(cls, frustum, voxel_coord, voxel_size, voxel_num, output_channels: int,
                     arith: Optional[int] = None, cache_plan: bool = False,
                     bev_channels_last: bool = False) -> "LiftSplat":
        """Build the view transform around the four buffers an existing ``LSSFPN`` / ``BSMLSSFPN`` already
        registered (lss_fpn.py:281-293), so that the very same fp32 values enter the kernels."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.is_bsm = False
        self.downsample_factor = None
        self.d_bound = None
        self.final_dim = None
        self.output_channels = int(output_channels)
        self.arith = arith
        self.cache_plan = cache_plan
        self.bev_channels_last = bool(bev_channels_last)
        self.register_buffer("voxel_size", voxel_size.detach().clone().float())
        self.register_buffer("voxel_coord", voxel_coord.detach().clone().float())
        self.register_buffer("voxel_num", voxel_num.detach().clone().long())
        self.register_buffer("frustum", frustum.detach().clone().float())
        self.height_channels = int(self.frustum.shape[0])
        self._grid = tuple(int(v) for v in self.voxel_num.tolist())
        self._plan_cache = {}
        self._grid_const = None
        return self

    def _const(self, device) -> _GridConst:
        gc = self._grid_const
        if gc is None or gc.device != device:
            gc = self._grid_const = _GridConst(self.frustum, self.voxel_coord, self.voxel_size, device)
        return gc

    # -- geometry -----------------------------------------------------------------------------------
    def get_geometry(self, sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat, reference_heights, bda_mat):
        """fp32 (B, Nc, D, fH, fW, 3) ego-frame points, as ``LSSFPN.get_geometry`` (lss_fpn.py:372-401)."""
        _, xyz = geometry_indices(self.frustum, sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat,
                                  reference_heights, bda_mat, self.voxel_coord, self.voxel_size, self.arith,
                                  return_xyz=True)
        return xyz

    def get_geometry_indices(self, sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat, reference_heights,
                             bda_mat):
        """int32 voxel indices of lss_fpn.py:487-488."""
        return geometry_indices(self.frustum, sensor2ego_mat, sensor2virtual_mat, intrin_mat, ida_mat,
                                reference_heights, bda_mat, self.voxel_coord, self.voxel_size, self.arith)

    def _auto_pipeline(self, frames: int, channels: int, inference: bool, plan_reused: bool) -> int:
        """AUTO policy, from the B200 measurements in profiles/README.md (round 2).  The voxel-tile pipeline is the faster
        one for training and for large inference batches; the pixel-block pipeline (one-kernel plan, no global sort) wins
        the small inference batches -- plan + forward 72 vs 97 us at one DAIR-R50 frame, 102 vs 189 us at one SGV3D-BSM-R50
        frame, break-even near 8 frames (16 for the dense stride-8 maps).  With a reused plan only the forward counts: the
        tile forward is faster on the stride-16 maps (28 vs 36 us), the block forward on the dense ones (62 vs 82 us)."""
        if _DEFAULT_PIPELINE != PIPELINE_AUTO:
            return _DEFAULT_PIPELINE
        d, fh, fw = (int(v) for v in self.frustum.shape[:3])
        # (256 x 256 and larger grids: footprints of up to 2 000 voxels per block are processed in rounds -- tile wins)
        if not inference or channels > 96 or d > 255 or self._grid[0] * self._grid[1] > 20000 or self.bev_channels_last:
            return PIPELINE_TILE
        dense = fh * fw >= 12000
        if plan_reused:
            return PIPELINE_BLOCK if (dense and frames <= 4) else PIPELINE_TILE
        return PIPELINE_BLOCK if frames <= (16 if dense else 8) else PIPELINE_TILE

    def make_plan(self, mats_dict, sweep_index: int = 0, channels: Optional[int] = None,
                  ctx_dtype: torch.dtype = torch.float32, inference: bool = False,
                  plan_reused: Optional[bool] = None) -> LiftSplatPlan:
        """``inference``: only the forward will run on this plan (lets AUTO pick the small-batch pipeline);
        ``plan_reused``: the plan outlives the step (defaults to ``cache_plan``)."""
        m = mats_dict
        args = (m["sensor2ego_mats"][:, sweep_index, ...], m["sensor2virtual_mats"][:, sweep_index, ...],
                m["intrin_mats"][:, sweep_index, ...], m["ida_mats"][:, sweep_index, ...],
                m["reference_heights"][:, sweep_index, ...], m.get("bda_mat", None))
        c = channels or self.output_channels
        reused = self.cache_plan if plan_reused is None else plan_reused
        pipe = self._auto_pipeline(int(args[0].shape[0]) * int(args[0].shape[1]), c, inference, reused)
        key = None
        # The cache is keyed on the identity of the calibration tensors (storage address, version counter, shape) AND
        # keeps those tensors alive next to the plan: while they live no other tensor can be allocated at their address,
        # and an in-place update bumps the version, so an equal key implies equal values.  Never used during stream
        # capture (a captured step must contain its own plan kernels).
        capturing = args[0].is_cuda and torch.cuda.is_current_stream_capturing()
        if self.cache_plan and not capturing:
            key = (c, ctx_dtype, pipe, self.bev_channels_last) + tuple(
                None if a is None else (a.data_ptr(), a._version, tuple(a.shape), tuple(a.stride()), a.dtype) for a in args)
            hit = self._plan_cache.get(key)
            if hit is not None:
                return hit[0]
        plan = LiftSplatPlan(self.frustum, *args, self.voxel_coord, self.voxel_size, self._grid, c, ctx_dtype,
                             self.arith, grid_const=self._const(args[0].device), pipeline=pipe,
                             channels_last=self.bev_channels_last and c == self.output_channels)
        if key is not None:
            self._plan_cache = {key: (plan, args)}
        return plan

    def invalidate_plan_cache(self) -> None:
        """Drop the cached plan (``cache_plan=True``), e.g. after re-calibrating a static camera out of band."""
        self._plan_cache = {}

    # -- call sites ---------------------------------------------------------------------------------
    def forward_single_sweep(self, height_feature: torch.Tensor, mats_dict, sweep_index: int = 0) -> torch.Tensor:
        """LSSFPN: ``height_feature`` = (B*Nc, D + C, fH, fW) output of the height net
        (lss_fpn.py:461).  Returns the (B, C, Y, X) contiguous BEV map of lss_fpn.py:494-495."""
        d, c = self.height_channels, self.output_channels
        hf = height_feature.float()
        inference = not (torch.is_grad_enabled() and hf.requires_grad)
        plan = self.make_plan(mats_dict, sweep_index, c, inference=inference)
        # softmax over the D logits (lss_fpn.py:462), lift (:464-466), geometry (:478-488) and pooling
        # (:490-495) all happen inside the library, on the head's output tensor in place
        return _LiftSplatHeadFunction.apply(hf, plan, d, c)

    def forward_single_sweep_bsm(self, height_logits, semantic_logits, context, mats_dict,
                                 sweep_index: int = 0) -> torch.Tensor:
        """BSMLSSFPN: ``out[0], out[1], out[2]`` of the MSCT head (bsm_lss_fpn.py:522-529): height
        logits (BN, D, fH, fW), 7 semantic logits, 80 context channels.  The 87-channel masked context of the
        reference (softmax, concat, background mask) is assembled inside the kernels, in the forward and -- under
        autograd -- in the backward as well (the torch assembly remains for the pixel-block pipeline, > 8 semantic
        channels, > 96 channels or non-fp32 context)."""
        if not (torch.is_grad_enabled() and (height_logits.requires_grad or semantic_logits.requires_grad
                                             or context.requires_grad)):
            # inference: softmax over the 7 semantic channels, concat and background mask (bsm_lss_fpn.py:524-529)
            # run inside the forward's context pass; the 87-channel tensor is never built
            plan = self.make_plan(mats_dict, sweep_index, int(context.shape[1] + semantic_logits.shape[1]), inference=True)
            return plan.forward_bsm(height_logits.float(), context.float(), semantic_logits.float(), 0.45)
        c_all = int(context.shape[1] + semantic_logits.shape[1])
        if c_all <= 96 and semantic_logits.shape[1] <= 8 and context.dtype == torch.float32:
            plan = self.make_plan(mats_dict, sweep_index, c_all)
            if not plan.uses_block_pipeline():
                # training: assembly and its backward fused into the kernels (no 87-channel tensor in either pass)
                return _LiftSplatBsmFunction.apply(height_logits.float(), semantic_logits.float(), context, plan, 0.45)
        semantic = semantic_logits.softmax(dim=1)                               # bsm_lss_fpn.py:524
        tran_feat = torch.cat((context, semantic), dim=1)                       # :526
        mask = semantic[:, 0, :, :].unsqueeze(1) > 0.45                         # :528 background
        tran_feat = tran_feat * (1 - mask.int())                                # :529
        plan = self.make_plan(mats_dict, sweep_index, int(tran_feat.shape[1]))
        # height softmax (bsm_lss_fpn.py:523) fused into the kernels
        return lift_splat(height_logits.float(), tran_feat.float(), plan, logits=True)


class LiftSplatGraph:
    """CUDA-graph replay of ``LiftSplat.forward_single_sweep`` for a serving loop with fixed shapes.

    The whole per-step sequence -- the reference's per-camera 4x4 products (lss_fpn.py:361,367,392), the
    plan kernels and the forward kernels -- is captured once and re-launched with a single
    ``cudaGraphLaunch``: the seven launches of a step otherwise cost about as much host time as the GPU
    needs to run them.  Inference only (no autograd).  The captured graph reads the tensors handed to
    the constructor (``height_feature`` and every entry of ``mats_dict``); ``__call__`` optionally
    copies new values into them first, and returns the (B, C, Y, X) BEV map, which is overwritten by
    the next call.  Geometry is recomputed on every replay, so the calibration may change per call (the plan
    cache of ``LiftSplat(cache_plan=True)`` is bypassed during capture).  Which ``mats_dict`` entries exist
    (``bda_mat`` present or ``None``), their shapes and dtypes are fixed at capture time; ``__call__`` raises on a mismatch.
    """

    def __init__(self, module: LiftSplat, height_feature: torch.Tensor, mats_dict, sweep_index: int = 0,
                 warmup: int = 2, static_calibration: bool = False):
        if not height_feature.is_cuda:
            raise RuntimeError("sgv3d_b200 runs on CUDA tensors only (no CPU fallback)")
        self.module, self.sweep_index = module, sweep_index
        self.height_feature = height_feature
        self.mats_dict = mats_dict
        self.device = height_feature.device
        self.static_calibration = static_calibration
        d, c = module.height_channels, module.output_channels
        # static roadside camera (IDA deterministic, BDA identity at inference: dataset/nusc_mv_det_dataset.py:433-454):
        # the voxel-run plan is built once, outside the graph; a replay is the two forward kernels only
        self.plan = module.make_plan(mats_dict, sweep_index, c, inference=True, plan_reused=True) if static_calibration \
            else None

        def step():
            if self.plan is not None:
                hf = height_feature.float()
                return self.plan.forward(hf[:, :d], hf[:, d:d + c], logits=True)
            return module.forward_single_sweep(height_feature, mats_dict, sweep_index)

        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):   # lazy initialisation (one-time self-checks, smem attributes) before capture
                step()
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.bev = step()

    def refresh_calibration(self, mats_dict=None) -> None:
        """``static_calibration`` mode: rebuild the plan in place (same workspace, so the captured graph
        stays valid) after the camera was re-calibrated."""
        if self.plan is None:
            raise RuntimeError("refresh_calibration() is for static_calibration=True graphs")
        m = mats_dict if mats_dict is not None else self.mats_dict
        i = self.sweep_index
        self.plan.update_calibration(m["sensor2ego_mats"][:, i, ...], m["sensor2virtual_mats"][:, i, ...],
                                     m["intrin_mats"][:, i, ...], m["ida_mats"][:, i, ...],
                                     m["reference_heights"][:, i, ...], m.get("bda_mat", None))

    def __call__(self, height_feature: Optional[torch.Tensor] = None, mats_dict=None) -> torch.Tensor:
        if height_feature is not None and height_feature is not self.height_feature:
            self.height_feature.copy_(height_feature, non_blocking=True)
        if mats_dict is not None and mats_dict is not self.mats_dict:
            if self.static_calibration:
                raise RuntimeError("static_calibration graph: call refresh_calibration(mats_dict) to change the matrices")
            for k, v in mats_dict.items():
                have = self.mats_dict.get(k)
                if v is None and have is None:
                    continue
                # whether an entry (e.g. bda_mat) is present is fixed at capture time: the graph has no buffer otherwise
                if (v is None) != (have is None):
                    raise RuntimeError(f"LiftSplatGraph: mats_dict[{k!r}] was {'absent' if have is None else 'present'} "
                                       f"at capture and cannot be {'added' if have is None else 'removed'} on replay")
                if tuple(v.shape) != tuple(have.shape) or v.dtype != have.dtype:
                    raise RuntimeError(f"LiftSplatGraph: mats_dict[{k!r}] is {tuple(v.shape)} {v.dtype}, captured "
                                       f"{tuple(have.shape)} {have.dtype}")
                have.copy_(v, non_blocking=True)
        self.graph.replay()
        return self.bev
