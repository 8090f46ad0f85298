"""Synthetic DAIR-V2X-I / Rope3D-shaped inputs for tests and ``bench.py``.

There is no dataset in the image, so calibrations are drawn from the distribution the
reference's data pipeline produces.  Citations are relative to /root/reference:

* ground plane in the camera frame / ``denorm``    dataset/nusc_mv_det_dataset.py:47-68
* ``sensor2virtual`` (Rodrigues, normal -> +y)      dataset/nusc_mv_det_dataset.py:70-82
* ``reference_heights`` (camera-to-plane distance)  dataset/nusc_mv_det_dataset.py:84-86
* IDA matrix (resize 0.8, zero crop, no flip/rot)   dataset/nusc_mv_det_dataset.py:133-161,433-446
* training-style intrinsic/extrinsic perturbation   dataset/nusc_mv_det_dataset.py:295-297,400-431
* Rope3D extrinsics (ego origin under the camera)   scripts/gen_info_rope3d.py:56-86

Everything here is plain numpy (no cv2): the values are *inputs* to the transform, so they only
need to be realistic, not bit-identical to the reference helpers.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .shapes import LiftSplatShape

__all__ = ["make_calibration", "make_mats", "make_activations", "ground_plane_in_camera",
           "sensor_to_virtual", "reference_height"]

SENSOR_H, SENSOR_W = 1080, 1920


def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], np.float64)


def _rot_z(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float64)


def ground_plane_in_camera(ego2sensor: np.ndarray) -> np.ndarray:
    """Plane ``ego z = 0`` expressed in the camera frame as (a, b, c, d), sign convention of the
    reference's ``get_denorm`` (three ego ground points -> camera, cross product, negated)."""
    pts = np.array([[0.0, 0.0, 0.0, 1.0], [0.0, 1.0, 0.0, 1.0], [1.0, 1.0, 0.0, 1.0]])
    cam = (ego2sensor.astype(np.float64) @ pts.T).T[:, :3]
    e1, e2 = cam[1] - cam[0], cam[2] - cam[0]
    n = np.cross(e1, e2)
    d = -float(n @ cam[0])
    return -1.0 * np.array([n[0], n[1], n[2], d])


def sensor_to_virtual(denorm: np.ndarray) -> np.ndarray:
    """4x4 fp32 rotation taking the downward ground normal onto camera +y (gravity-aligned
    "virtual" camera), via the axis-angle (Rodrigues) formula."""
    target = -1.0 * denorm[:3]
    target = target / np.linalg.norm(target)
    origin = np.array([0.0, 1.0, 0.0])
    angle = math.acos(max(-1.0, min(1.0, float(target @ origin))))
    axis = np.cross(target, origin)
    norm = np.linalg.norm(axis)
    out = np.eye(4, dtype=np.float32)
    if norm < 1e-12:
        return out
    k = (axis / norm).astype(np.float32).astype(np.float64)
    kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    rot = np.eye(3) + math.sin(angle) * kx + (1.0 - math.cos(angle)) * (kx @ kx)
    out[:3, :3] = rot.astype(np.float32)
    return out


def reference_height(denorm: np.ndarray) -> np.float32:
    return np.float32(abs(denorm[3]) / np.linalg.norm(denorm[:3]))


def make_calibration(rng: np.random.Generator, family: str = "dair", final_dim=(864, 1536),
                     perturb: bool = False) -> Dict[str, np.ndarray]:
    """One camera: dict of fp32 4x4 ``sensor2ego``, ``sensor2virtual``, ``intrin``, ``ida`` and the
    scalar ``reference_height``.  ``family``: "dair" (lidar-frame ego, camera 5-8 m above the ego
    origin plane) or "rope3d" (ego origin on the ground under the camera)."""
    fx, fy = rng.uniform(2100.0, 2400.0, 2)
    cx = SENSOR_W / 2 + rng.uniform(-30.0, 30.0)
    cy = SENSOR_H / 2 + rng.uniform(-30.0, 30.0)
    cam_h = rng.uniform(5.0, 8.0)
    pitch = math.radians(rng.uniform(8.0, 14.0))
    roll = math.radians(rng.normal(0.0, 1.0))
    if perturb:  # dataset/nusc_mv_det_dataset.py:400-431
        ratio = rng.normal(1.0, 0.2)
        fx, fy = fx * ratio, fy * ratio
        roll += math.radians(rng.normal(0.0, 2.0))
        pitch += math.radians(rng.normal(0.0, 0.67))
    # camera axes (x right, y down, z forward) in an ego frame with x forward, y left, z up
    base = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    rot = base @ _rot_x(-pitch) @ _rot_z(roll)
    sensor2ego = np.eye(4)
    sensor2ego[:3, :3] = rot
    if family == "rope3d":
        sensor2ego[:3, 3] = [0.0, 0.0, cam_h]
    else:
        sensor2ego[:3, 3] = [rng.uniform(-1.0, 1.0), rng.uniform(-2.0, 2.0), cam_h]
    sensor2ego = sensor2ego.astype(np.float32)
    ego2sensor = np.linalg.inv(sensor2ego.astype(np.float64))
    denorm = ground_plane_in_camera(ego2sensor)
    intrin = np.zeros((4, 4), np.float32)
    intrin[0, 0], intrin[1, 1], intrin[0, 2], intrin[1, 2] = fx, fy, cx, cy
    intrin[2, 2] = intrin[3, 3] = 1.0
    # IDA: resize = max(fH/H, fW/W), zero crop / flip / rotation (dataset/...:433-446,133-161)
    resize = max(final_dim[0] / SENSOR_H, final_dim[1] / SENSOR_W)
    ida = np.eye(4, dtype=np.float32)
    ida[0, 0] = ida[1, 1] = np.float32(resize)
    return dict(sensor2ego=sensor2ego, sensor2virtual=sensor_to_virtual(denorm), intrin=intrin,
                ida=ida, reference_height=reference_height(denorm))


def make_mats(shape: LiftSplatShape, batch: int, num_cams: int = 1, seed: int = 0,
              perturb_fraction: float = 0.5, device="cpu", bda: Optional[str] = "identity"
              ) -> Dict[str, torch.Tensor]:
    """``mats_dict``-like tensors for one sweep: sensor2ego / sensor2virtual / intrin / ida
    (B, Nc, 4, 4), reference_heights (B, Nc), bda (B, 4, 4) or None -- the slices
    ``_forward_single_sweep`` passes to ``get_geometry`` (layers/backbones/lss_fpn.py:478-485)."""
    rng = np.random.default_rng(seed)
    keys = ("sensor2ego", "sensor2virtual", "intrin", "ida")
    out = {k: np.zeros((batch, num_cams, 4, 4), np.float32) for k in keys}
    ref_h = np.zeros((batch, num_cams), np.float32)
    for b in range(batch):
        for n in range(num_cams):
            cal = make_calibration(rng, shape.family, shape.final_dim,
                                   perturb=bool(rng.random() < perturb_fraction))
            for k in keys:
                out[k][b, n] = cal[k]
            ref_h[b, n] = cal["reference_height"]
    mats = {k: torch.from_numpy(v).to(device) for k, v in out.items()}
    mats["reference_heights"] = torch.from_numpy(ref_h).to(device)
    if bda == "identity":
        mats["bda"] = torch.eye(4).repeat(batch, 1, 1).to(device)
    elif bda == "random":
        mm = np.tile(np.eye(4, dtype=np.float32), (batch, 1, 1))
        for b in range(batch):
            a = math.radians(rng.uniform(-5.0, 5.0))
            s = rng.uniform(0.95, 1.05)
            mm[b, :3, :3] = (_rot_z(a) * s).astype(np.float32)
        mats["bda"] = torch.from_numpy(mm).to(device)
    else:
        mats["bda"] = None
    return mats


def make_activations(shape: LiftSplatShape, batch: int, num_cams: int = 1, seed: int = 0,
                     device="cpu", peaky: bool = False, channels: Optional[int] = None,
                     generator_device: Optional[str] = None):
    """Height logits (B*Nc, D, fH, fW) ~ N(0,1) (x4 if ``peaky``) and context (B*Nc, C, fH, fW) ~ N(0,1)."""
    c = channels or shape.channels
    gdev = generator_device or device
    gen = torch.Generator(device=gdev)
    gen.manual_seed(1234 + seed)
    bn = batch * num_cams
    logits = torch.randn(bn, shape.D, shape.fH, shape.fW, generator=gen, device=gdev)
    if peaky:
        logits = logits * 4.0
    ctx = torch.randn(bn, c, shape.fH, shape.fW, generator=gen, device=gdev)
    return logits.to(device), ctx.to(device)
