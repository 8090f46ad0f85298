#!/usr/bin/env python
"""BSMLSSFPN call site (bsm_lss_fpn.py:523-559) at the SGV3D-BSM-R50 shape: fused context assembly vs the torch
assembly followed by the plain forward; per-step device time (plan rebuilt every step)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgv3d_b200 import LiftSplat, get_shape  # noqa: E402
from sgv3d_b200.synthetic import make_mats  # noqa: E402
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=8); ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--background", type=float, default=-1.0, help="fraction of background pixels (semantic[0] > 0.45); default: random logits")
ap.add_argument("--pipeline", default="auto", choices=["auto", "tile", "block"])
a = ap.parse_args()
from sgv3d_b200 import view_transform as VT  # noqa: E402
VT.set_default_pipeline({"auto": VT.PIPELINE_AUTO, "tile": VT.PIPELINE_TILE, "block": VT.PIPELINE_BLOCK}[a.pipeline])
s = get_shape("sgv3d_bsm_r50"); dev = torch.device("cuda", 0); B = a.batch
mod = LiftSplat(s.x_bound, s.y_bound, s.z_bound, s.d_bound, s.final_dim, s.downsample, 87).to(dev)
mats = make_mats(s, B, 1, seed=5, bda="identity")
md = {"sensor2ego_mats": mats["sensor2ego"].unsqueeze(1).to(dev), "sensor2virtual_mats": mats["sensor2virtual"].unsqueeze(1).to(dev),
      "intrin_mats": mats["intrin"].unsqueeze(1).to(dev), "ida_mats": mats["ida"].unsqueeze(1).to(dev),
      "reference_heights": mats["reference_heights"].unsqueeze(1).to(dev), "bda_mat": mats["bda"].to(dev)}
hl = torch.randn(B, s.D, s.fH, s.fW, device=dev); sl = torch.randn(B, 7, s.fH, s.fW, device=dev); cx = torch.randn(B, 80, s.fH, s.fW, device=dev)
if a.background >= 0:   # real roadside frames are mostly background: push channel 0 up on that fraction of the pixels
    bg = torch.rand(B, s.fH, s.fW, device=dev) < a.background
    sl[:, 0] = torch.where(bg, torch.full_like(sl[:, 0], 6.0), torch.full_like(sl[:, 0], -6.0))
print("background fraction", float((sl.softmax(1)[:, 0] > 0.45).float().mean()), "pipeline", a.pipeline, flush=True)

def fused():
    with torch.no_grad():
        return mod.forward_single_sweep_bsm(hl, sl, cx, md)

def torch_assembly():
    with torch.no_grad():
        semantic = sl.softmax(dim=1)
        tf = torch.cat((cx, semantic), dim=1)
        mask = semantic[:, 0, :, :].unsqueeze(1) > 0.45
        tf = tf * (1 - mask.int())
        return mod.make_plan(md, 0, 87).forward(hl, tf, logits=True)

for name, fn in (("torch assembly + forward", torch_assembly), ("fused assembly", fused)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(f"sgv3d_bsm_r50 batch {B} {name}: {1e3 * ms:.1f} us/step  {B / ms * 1e3:.0f} frames/s", flush=True)
