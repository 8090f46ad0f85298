#!/usr/bin/env python
"""Tiny driver for ncu: runs a few lift-splat steps (plan + forward [+ backward]) at a given batch."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgv3d_b200 import LiftSplat, get_shape  # noqa: E402
from sgv3d_b200.synthetic import make_activations, make_mats  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="dair_r50")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--backward", action="store_true")
a = ap.parse_args()
s = get_shape(a.shape)
dev = torch.device("cuda", 0)
mod = LiftSplat(s.x_bound, s.y_bound, s.z_bound, s.d_bound, s.final_dim, s.downsample, s.channels).to(dev)
mats = make_mats(s, a.batch, 1, seed=5, bda="identity")
md = {"sensor2ego_mats": mats["sensor2ego"].unsqueeze(1).to(dev), "sensor2virtual_mats": mats["sensor2virtual"].unsqueeze(1).to(dev),
      "intrin_mats": mats["intrin"].unsqueeze(1).to(dev), "ida_mats": mats["ida"].unsqueeze(1).to(dev),
      "reference_heights": mats["reference_heights"].unsqueeze(1).to(dev), "bda_mat": mats["bda"].to(dev)}
logits, ctx = make_activations(s, a.batch, 1, seed=5, device=dev, generator_device=dev)
hf = torch.cat((logits, ctx), 1).contiguous().requires_grad_(a.backward)
for _ in range(a.steps):
    bev = mod.forward_single_sweep(hf, md)
    if a.backward:
        bev.backward(torch.ones_like(bev))
torch.cuda.synchronize()
print("done", tuple(bev.shape))
