"""Shared helpers for the test-suite (loading golden fixtures, frustum axes, etc.)."""
import glob
import os

import numpy as np
import torch

from oracle import lift_splat_oracle as O
from sgv3d_b200.shapes import get_shape

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[len("golden_"):-len(".npz")]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "golden_*.npz")))


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"golden_{name}.npz"), allow_pickle=False))
    g["shape"] = get_shape(str(g["shape"]))
    g["batch"], g["num_cams"] = int(g["batch"]), int(g["num_cams"])
    if g["bda"].size == 0:
        g["bda"] = None
    return g


def golden_mats(g, device="cpu"):
    m = {k: torch.from_numpy(g[k]).to(device) for k in
         ("sensor2ego", "sensor2virtual", "intrin", "ida", "reference_heights")}
    m["bda"] = torch.from_numpy(g["bda"]).to(device) if g["bda"] is not None else None
    return m


def frustum_axes(frustum):
    """(u[fW], v[fH], z[D]) sliced out of a (D,fH,fW,4) frustum buffer."""
    return frustum[0, 0, :, 0].contiguous(), frustum[0, :, 0, 1].contiguous(), frustum[:, 0, 0, 2].contiguous()


def kept_mask_np(idx, grid):
    X, Y, Z = grid
    return ((idx[..., 0] >= 0) & (idx[..., 0] < X) & (idx[..., 1] >= 0) & (idx[..., 1] < Y)
            & (idx[..., 2] >= 0) & (idx[..., 2] < Z))


def oracle_frustum(shape):
    return O.create_frustum(shape.final_dim, shape.downsample, shape.d_bound)
