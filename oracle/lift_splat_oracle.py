"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference lift-splat path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import this module.  The product package ``sgv3d_b200`` never does.

Parity pin: the reference (yanglei18/SGV3D) ships no tests or golden vectors for this path
(SURVEY.md §4, §8c).  This port is pinned against the reference's own Python, imported
unmodified in the build container by ``tests/golden/make_golden.py``; the outputs are
committed under ``tests/golden/`` and ``tests/test_oracle_golden.py`` replays them.

All citations are relative to ``/root/reference``.  Two flavours of the geometry exist:

* ``geometry_matmul``   -- same torch calls, same broadcast shapes as the reference
  (``get_geometry`` lss_fpn.py:372-401, ``height2localtion`` lss_fpn.py:350-370).  Runs on
  any torch device; this is what ``bench.py`` times as the CPU baseline, and what tests run on
  ``cuda`` to learn the order cuBLAS evaluates the 4-term dot products in.
* ``geometry_explicit`` -- numpy fp32, every rounding spelled out, torch-CPU order
  ``((a0*b0 + a1*b1) + a2*b2) + a3*b3`` (SURVEY.md §7 hard part 1).  The C oracle
  (``oracle/sgv3d_oracle.c``) implements the same plus the FMA-chain order.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = [
    "create_frustum", "grid_buffers", "camera_matrices", "geometry_matmul", "geometry_explicit",
    "quantize", "quantize_np", "voxel_pooling_forward", "voxel_pooling_backward", "lift",
    "bsm_context", "lift_splat_forward", "lift_splat_forward_backward",
]


# --------------------------------------------------------------------------------------
# module buffers
# --------------------------------------------------------------------------------------
def create_frustum(final_dim, downsample_factor, d_bound) -> torch.Tensor:
    """(D, fH, fW, 4) fp32 frustum buffer; restates ``LSSFPN.create_frustum``
    (layers/backbones/lss_fpn.py:325-348; bsm_lss_fpn.py:384-407).

    u = linspace(0, W_in-1, fW), v = linspace(0, H_in-1, fH) (torch fp32 linspace);
    z_d = d0 + (d/D)**1.5 * (d1-d0) evaluated in float64 numpy then cast to fp32 ("DID").
    """
    in_h, in_w = final_dim
    f_h, f_w = in_h // downsample_factor, in_w // downsample_factor
    n_bins = d_bound[2]
    frac = np.power(np.arange(n_bins) / n_bins, 1.5)
    z64 = d_bound[0] + frac * (d_bound[1] - d_bound[0])
    z = torch.tensor(z64, dtype=torch.float)
    u = torch.linspace(0, in_w - 1, f_w, dtype=torch.float)
    v = torch.linspace(0, in_h - 1, f_h, dtype=torch.float)
    out = torch.empty(n_bins, f_h, f_w, 4, dtype=torch.float)
    out[..., 0] = u.view(1, 1, f_w)
    out[..., 1] = v.view(1, f_h, 1)
    out[..., 2] = z.view(n_bins, 1, 1)
    out[..., 3] = 1.0
    return out


def grid_buffers(x_bound, y_bound, z_bound):
    """voxel_size fp32[3], voxel_coord fp32[3], voxel_num int64[3]; restates the
    ``register_buffer`` calls at lss_fpn.py:281-292 (bsm_lss_fpn.py:349-360)."""
    rows = [x_bound, y_bound, z_bound]
    voxel_size = torch.Tensor([r[2] for r in rows])
    voxel_coord = torch.Tensor([r[0] + r[2] / 2.0 for r in rows])
    voxel_num = torch.LongTensor([(r[1] - r[0]) / r[2] for r in rows])
    return voxel_size, voxel_coord, voxel_num


# --------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------
def camera_matrices(sensor2ego, sensor2virtual, intrin, ida):
    """Per-camera 4x4 products with the SAME torch calls the reference makes, so the 16-float
    operands of the per-point math are identical (SURVEY.md §7 hard part 1):
    ``ida.inverse()`` (lss_fpn.py:392), ``sensor2virtual.matmul(torch.inverse(intrin))`` (:361),
    ``sensor2ego.matmul(torch.inverse(sensor2virtual))`` (:367).  Shapes (B, Nc, 4, 4)."""
    ida_inv = ida.inverse()
    m_virtual = sensor2virtual.matmul(torch.inverse(intrin))
    m_ego = sensor2ego.matmul(torch.inverse(sensor2virtual))
    return ida_inv, m_virtual, m_ego


def geometry_matmul(frustum, sensor2ego, sensor2virtual, intrin, ida, ref_heights, bda):
    """(B, Nc, D, fH, fW, 3) fp32 ego-frame points through broadcast ``matmul`` exactly like
    ``get_geometry`` (lss_fpn.py:372-401) + ``height2localtion`` (lss_fpn.py:350-370)."""
    b, nc = sensor2ego.shape[:2]
    d, fh, fw, _ = frustum.shape
    ida_inv, m_virtual, m_ego = camera_matrices(sensor2ego, sensor2virtual, intrin, ida)
    # :391-392 undo image-data augmentation
    pts = ida_inv.view(b, nc, 1, 1, 1, 4, 4).matmul(frustum.unsqueeze(-1))
    # :352-354 height of the camera above the plane z = z_d
    rh = ref_heights.view(b, nc, 1, 1, 1, 1).expand(b, nc, d, fh, fw, 1)
    hgt = -1 * pts[..., 2, :] + rh
    # :356-360 pixel ray at virtual depth 10: (10x, 10y, 10, w)
    ray = pts.clone()
    ray[..., 2, :] = 10
    ray = torch.cat((ray[..., :2, :] * ray[..., 2:3, :], ray[..., 2:, :]), dim=-2)
    # :361-362 into the gravity-aligned virtual camera
    pv = m_virtual.view(b, nc, 1, 1, 1, 4, 4).matmul(ray)
    # :363-366 scale the ray so that its drop equals the height; homogeneous coord := 1
    ratio = hgt[..., 0] / pv[..., 1, 0]
    pe = pv * ratio.view(b, nc, d, fh, fw, 1, 1)
    pe[..., 3, :] = 1
    # :367-369 virtual camera -> ego
    pg = m_ego.view(b, nc, 1, 1, 1, 4, 4).matmul(pe)
    # :394-398 bev data augmentation
    if bda is not None:
        pg = bda.view(b, 1, 1, 1, 1, 4, 4).expand(b, nc, 1, 1, 1, 4, 4) @ pg
    return pg.squeeze(-1)[..., :3]


def _dot4_seq(m_row, v0, v1, v2, v3):
    """((m0*v0 + m1*v1) + m2*v2) + m3*v3 with fp32 rounding after every op."""
    acc = m_row[..., 0] * v0
    acc = acc + m_row[..., 1] * v1
    acc = acc + m_row[..., 2] * v2
    acc = acc + m_row[..., 3] * v3
    return acc


def geometry_explicit(u_tab, v_tab, z_tab, ida_inv, m_virtual, m_ego, ref_heights, bda):
    """numpy-fp32 restatement of the per-point math, torch-CPU
    evaluation order.  ``*_tab`` are the frustum axes; matrices are (B, Nc, 4, 4) fp32 numpy,
    ``bda`` (B, 4, 4) or None.  Returns (B, Nc, D, fH, fW, 3) fp32."""
    f32 = np.float32
    u = np.asarray(u_tab, f32).reshape(1, 1, 1, 1, -1)
    v = np.asarray(v_tab, f32).reshape(1, 1, 1, -1, 1)
    z = np.asarray(z_tab, f32).reshape(1, 1, -1, 1, 1)
    one = f32(1.0)

    def rows(m):  # (B,Nc,4,4) -> list of 4 broadcastable row views (B,Nc,1,1,1,4)
        m = np.asarray(m, f32)
        return [m[:, :, r].reshape(m.shape[0], m.shape[1], 1, 1, 1, 4) for r in range(4)]

    a, mv, me = rows(ida_inv), rows(m_virtual), rows(m_ego)
    p0 = [_dot4_seq(a[r], u, v, z, one) for r in range(4)]              # lss_fpn.py:392
    rh = np.asarray(ref_heights, f32).reshape(ref_heights.shape[0], -1, 1, 1, 1)
    hgt = (f32(-1.0) * p0[2]) + rh                                        # :354
    q0, q1, q2, q3 = p0[0] * f32(10.0), p0[1] * f32(10.0), f32(10.0), p0[3]  # :356-360
    pv = [_dot4_seq(mv[r], q0, q1, q2, q3) for r in range(4)]             # :361-362
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = hgt / pv[1]                                               # :363
        e0, e1, e2 = pv[0] * ratio, pv[1] * ratio, pv[2] * ratio          # :365
    pg = [_dot4_seq(me[r], e0, e1, e2, one) for r in range(4)]            # :366-369
    if bda is not None:                                                   # :394-398
        bm = np.asarray(bda, f32)
        br = [bm[:, r].reshape(bm.shape[0], 1, 1, 1, 1, 4) for r in range(3)]
        with np.errstate(invalid="ignore"):
            pg = [_dot4_seq(br[r], pg[0], pg[1], pg[2], pg[3]) for r in range(3)]
    shape = np.broadcast_shapes(*(np.shape(c) for c in pg[:3]))
    return np.stack([np.broadcast_to(c, shape) for c in pg[:3]], axis=-1).astype(f32)


def quantize(geom, voxel_coord, voxel_size):
    """int32 voxel indices, restating lss_fpn.py:487-488 with torch ops.  NB on CPU torch's
    float->int cast of NaN/out-of-range values is x86 ``cvttss2si`` (INT_MIN); on CUDA it
    saturates and maps NaN to 0 (SURVEY.md §7 hard part 2).  ``quantize_np`` has CUDA semantics."""
    return ((geom - (voxel_coord - voxel_size / 2.0)) / voxel_size).int()


def quantize_np(geom, voxel_coord, voxel_size):
    """numpy version of ``quantize`` with the GPU's ``cvt.rzi.s32.f32`` semantics
    (truncate toward zero, saturate, NaN -> 0) -- what the reference produces in production,
    where the op only exists on CUDA."""
    f32 = np.float32
    lower = (np.asarray(voxel_coord, f32) - np.asarray(voxel_size, f32) / f32(2.0)).astype(f32)
    with np.errstate(invalid="ignore", over="ignore"):
        q = ((np.asarray(geom, f32) - lower) / np.asarray(voxel_size, f32)).astype(f32)
        t = np.trunc(q.astype(np.float64))
    t = np.where(np.isnan(t), 0.0, t)
    t = np.clip(t, -2147483648.0, 2147483647.0)
    return t.astype(np.int64).astype(np.int32)


# --------------------------------------------------------------------------------------
# voxel pooling (the op)
# --------------------------------------------------------------------------------------
def voxel_pooling_forward(geom_xyz, input_features, voxel_num, accumulate_dtype=None):
    """torch-CPU ``index_add_`` equivalent of the CUDA-only op: restates
    ``VoxelPooling.forward`` (ops/voxel_pooling/voxel_pooling.py:9-55) and the kernel
    (ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:16-34).  Returns
    ``(bev (B,C,Y,X) permuted view, pos_memo (B,N,3) int32)``."""
    b = geom_xyz.shape[0]
    geom = geom_xyz.reshape(b, -1, 3)
    feat = input_features.reshape(b, -1, input_features.shape[-1])
    assert geom.shape[1] == feat.shape[1]
    n, c = feat.shape[1], feat.shape[2]
    nx, ny, nz = (int(voxel_num[0]), int(voxel_num[1]), int(voxel_num[2]))
    x, y, z = geom[..., 0].long(), geom[..., 1].long(), geom[..., 2].long()
    kept = (x >= 0) & (x < nx) & (y >= 0) & (y < ny) & (z >= 0) & (z < nz)      # .cu:24
    batch = torch.arange(b, device=geom.device).view(b, 1).expand(b, n)
    pos_memo = geom.new_full((b, n, 3), -1)                                     # .py:40
    pos_memo[kept] = torch.stack((batch[kept], y[kept], x[kept]), -1).to(pos_memo.dtype)  # .cu:27-29
    lin = (batch * ny + y) * nx + x                                             # .cu:32
    acc_dtype = accumulate_dtype or feat.dtype
    out = torch.zeros(b * ny * nx, c, dtype=acc_dtype, device=feat.device)      # .py:37-38
    out.index_add_(0, lin[kept], feat[kept].to(acc_dtype))                      # .cu:30-34
    return out.view(b, ny, nx, c).permute(0, 3, 1, 2), pos_memo                 # .py:55


def voxel_pooling_backward(grad_output, pos_memo, num_channels):
    """restates ``VoxelPooling.backward`` (ops/voxel_pooling/voxel_pooling.py:57-69):
    ``grad_feat[kept] = grad_out[b, :, y, x]``, zero elsewhere.  Returns (B, N, C)."""
    b, n, _ = pos_memo.shape
    kept = pos_memo[..., 0] != -1
    grad = grad_output.new_zeros(b, n, num_channels)
    pm = pos_memo[kept].long()
    grad[kept] = grad_output[pm[:, 0], :, pm[:, 1], pm[:, 2]]
    return grad


# --------------------------------------------------------------------------------------
# call-site glue
# --------------------------------------------------------------------------------------
def lift(height_prob, context):
    """Outer product "lift": (BN, D, fH, fW) x (BN, C, fH, fW) -> (BN, C, D, fH, fW);
    restates lss_fpn.py:464-466 (bsm_lss_fpn.py:531)."""
    return height_prob.unsqueeze(1) * context.unsqueeze(2)


def bsm_context(context, semantic_logits, threshold=0.45):
    """BSM context assembly, restating bsm_lss_fpn.py:524-529: 7-way semantic softmax,
    concat behind the 80 context channels, zero the pixels whose background probability
    exceeds 0.45."""
    semantic = semantic_logits.softmax(dim=1)
    feat = torch.cat((context, semantic), dim=1)
    mask = semantic[:, 0, :, :].unsqueeze(1) > threshold
    return feat * (1 - mask.int())


def lift_splat_forward(height_logits, context, frustum, mats, voxel_coord, voxel_size, voxel_num,
                       geometry=geometry_matmul):
    """Whole reference path for one sweep, restating ``_forward_single_sweep``
    lss_fpn.py:462-495: softmax over D, lift, geometry, permute, quantise, voxel pooling,
    final ``.contiguous()``.  ``mats`` holds sensor2ego/sensor2virtual/intrin/ida (B,Nc,4,4),
    reference_heights (B,Nc), bda (B,4,4) or None.  Returns (bev (B,C,Y,X), idx, pos_memo)."""
    b, nc = mats["sensor2ego"].shape[:2]
    height = height_logits.softmax(1)                                            # :462
    feat = lift(height, context)                                                 # :464-466
    feat = feat.reshape(b, nc, *feat.shape[1:])                                  # :469-476
    geom = geometry(frustum, mats["sensor2ego"], mats["sensor2virtual"], mats["intrin"],
                    mats["ida"], mats["reference_heights"], mats.get("bda"))     # :478-485
    feat = feat.permute(0, 1, 3, 4, 5, 2)                                        # :486
    idx = quantize(geom, voxel_coord, voxel_size)                                # :487-488
    bev, pos_memo = voxel_pooling_forward(idx, feat.contiguous(), voxel_num)     # :490-491
    return bev.contiguous(), idx, pos_memo                                       # :494-495


def lift_splat_forward_backward(height_logits, context, frustum, mats, voxel_coord, voxel_size,
                                voxel_num, grad_bev):
    """Forward + gradients w.r.t. the height logits and the context, through torch autograd
    over the port above (index_add_'s autograd is the gather of voxel_pooling.py:57-69)."""
    hl = height_logits.detach().clone().requires_grad_(True)
    cx = context.detach().clone().requires_grad_(True)
    b, nc = mats["sensor2ego"].shape[:2]
    height = hl.softmax(1)
    feat = lift(height, cx)
    feat = feat.reshape(b, nc, *feat.shape[1:]).permute(0, 1, 3, 4, 5, 2)
    with torch.no_grad():
        geom = geometry_matmul(frustum, mats["sensor2ego"], mats["sensor2virtual"], mats["intrin"],
                               mats["ida"], mats["reference_heights"], mats.get("bda"))
        idx = quantize(geom, voxel_coord, voxel_size)
    bev, _ = voxel_pooling_forward(idx, feat.contiguous(), voxel_num)
    bev = bev.contiguous()
    bev.backward(grad_bev)
    return bev.detach(), hl.grad, cx.grad
