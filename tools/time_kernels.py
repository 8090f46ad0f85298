#!/usr/bin/env python
"""Per-kernel CUDA-event timings (the library's own profiler) of plan / forward / backward at a given batch."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgv3d_b200 import LiftSplat, get_shape, _native as N  # noqa: E402
from sgv3d_b200.synthetic import make_activations, make_mats  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="dair_r50"); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--iters", type=int, default=20); ap.add_argument("--bf16", action="store_true")
ap.add_argument("--pipeline", default="auto", choices=["auto", "tile", "block"])
ap.add_argument("--bda", default="identity", choices=["identity", "random", "none"])
ap.add_argument("--channels-last", action="store_true", help="BEV map / gradient in torch.channels_last order")
a = ap.parse_args()
from sgv3d_b200 import view_transform as VT  # noqa: E402
VT.set_default_pipeline({"auto": VT.PIPELINE_AUTO, "tile": VT.PIPELINE_TILE, "block": VT.PIPELINE_BLOCK}[a.pipeline])
s = get_shape(a.shape); dev = torch.device("cuda", 0)
mod = LiftSplat(s.x_bound, s.y_bound, s.z_bound, s.d_bound, s.final_dim, s.downsample, s.channels,
                bev_channels_last=a.channels_last).to(dev)
mats = make_mats(s, a.batch, 1, seed=5, bda=None if a.bda == "none" else a.bda)
md = {"sensor2ego_mats": mats["sensor2ego"].unsqueeze(1).to(dev), "sensor2virtual_mats": mats["sensor2virtual"].unsqueeze(1).to(dev),
      "intrin_mats": mats["intrin"].unsqueeze(1).to(dev), "ida_mats": mats["ida"].unsqueeze(1).to(dev),
      "reference_heights": mats["reference_heights"].unsqueeze(1).to(dev),
      "bda_mat": mats["bda"].to(dev) if mats["bda"] is not None else None}
logits, ctx = make_activations(s, a.batch, 1, seed=5, device=dev, generator_device=dev)
if a.bf16:
    ctx = ctx.bfloat16()
plan = mod.make_plan(md, 0, s.channels, ctx.dtype)
gb = torch.randn(a.batch, s.channels, s.grid[1], s.grid[0], device=dev)
if a.channels_last:
    gb = gb.contiguous(memory_format=torch.channels_last)
for phase, fn in (("plan", plan.rebuild), ("forward", lambda: plan.forward(logits, ctx, logits=True)),
                  ("backward", lambda: plan.backward(gb, logits, ctx, logits=True))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    N.profile_enable(True); N.profile_report()
    for _ in range(a.iters):
        fn()
    torch.cuda.synchronize()
    prof = N.profile_report(); N.profile_enable(False)
    tot = sum(t for _, t in prof.values()) / a.iters
    print(f"[{a.shape} B={a.batch} {a.pipeline}] {phase}: {1e3 * tot:.1f} us  " + "  ".join(f"{k}={1e3 * t / n:.1f}" for k, (n, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])))
