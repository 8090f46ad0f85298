#!/usr/bin/env python
"""BASELINE configs 3-5: fused lift-splat over shapes / height bins / channels / grids / context dtype.
One JSON line per case: device time of plan, forward (cached plan) and backward, frames/s of the step
(plan + forward) and of the training step (plan + forward + backward), achieved algorithmic GB/s.
    python tools/sweep.py [--batch 8] [--iters 10] > profiles/sweep_r01.jsonl"""
import argparse, dataclasses, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgv3d_b200 import LiftSplat, get_shape  # noqa: E402
from sgv3d_b200.synthetic import make_activations, make_mats  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8); ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda", 0)
PEAK = 6553.0
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def case(tag, s, B, ctx_dtype=torch.float32):
    mod = LiftSplat(s.x_bound, s.y_bound, s.z_bound, s.d_bound, s.final_dim, s.downsample, s.channels).to(dev)
    mats = make_mats(s, B, 1, seed=5, bda="identity")
    md = {"sensor2ego_mats": mats["sensor2ego"].unsqueeze(1).to(dev), "sensor2virtual_mats": mats["sensor2virtual"].unsqueeze(1).to(dev),
          "intrin_mats": mats["intrin"].unsqueeze(1).to(dev), "ida_mats": mats["ida"].unsqueeze(1).to(dev),
          "reference_heights": mats["reference_heights"].unsqueeze(1).to(dev), "bda_mat": mats["bda"].to(dev)}
    logits, ctx = make_activations(s, B, 1, seed=5, device=dev, generator_device=dev)
    ctx = ctx.to(ctx_dtype)
    plan = mod.make_plan(md, 0, s.channels, ctx_dtype)
    gb = torch.randn(B, s.channels, s.grid[1], s.grid[0], device=dev)
    ms_plan = timeit(plan.rebuild, a.iters)
    ms_fwd = timeit(lambda: plan.forward(logits, ctx, logits=True), a.iters)
    ms_bwd = timeit(lambda: plan.backward(gb, logits, ctx, logits=True), a.iters)
    cb = 2 if ctx_dtype == torch.bfloat16 else 4
    fb, bb = s.fused_forward_bytes(cb) * B, s.fused_backward_bytes(cb) * B
    out = {"case": tag, "shape": s.name, "frames": B, "D": s.D, "fH": s.fH, "fW": s.fW, "C": s.channels, "grid": list(s.grid),
           "ctx": "bf16" if cb == 2 else "f32", "us_plan": 1e3 * ms_plan, "us_forward": 1e3 * ms_fwd, "us_backward": 1e3 * ms_bwd,
           "step_frames_per_s": B / ((ms_plan + ms_fwd) * 1e-3), "train_frames_per_s": B / ((ms_plan + ms_fwd + ms_bwd) * 1e-3),
           "forward_cached_plan_GBs": fb / ms_fwd / 1e6, "forward_cached_plan_frac": fb / ms_fwd / 1e6 / PEAK,
           "step_GBs": fb / (ms_plan + ms_fwd) / 1e6, "train_GBs": (fb + bb) / (ms_plan + ms_fwd + ms_bwd) / 1e6,
           "train_frac": (fb + bb) / (ms_plan + ms_fwd + ms_bwd) / 1e6 / PEAK}
    print(json.dumps(out), flush=True)
    del plan, mod
    torch.cuda.empty_cache()


B = a.batch
# config 3 / appendix B: the named shapes
for name in ("dair_r50", "dair_r50_256", "rope3d_r50", "rope3d_r101_256", "rope3d_r101_140", "rope3d_native", "sgv3d_bsm_r50",
             "sgv3d_bsm_r101"):
    case("shape", get_shape(name), B if "bsm_r101" not in name else max(1, B // 2))
# config 4: bf16 context
for name in ("dair_r50", "sgv3d_bsm_r50"):
    case("bf16-context", get_shape(name), B, torch.bfloat16)
# config 5: D x C x grid microbench sweep at 54 x 96
base = get_shape("rope3d_r50")
for D in (60, 90, 120, 180):
    for C in (64, 80, 87, 128, 256):
        for g in ("128", "256"):
            gs = dict(x_bound=(0.0, 102.4, 0.8), y_bound=(-51.2, 51.2, 0.8)) if g == "128" else \
                dict(x_bound=(0.0, 102.4, 0.4), y_bound=(-51.2, 51.2, 0.4))
            s = dataclasses.replace(base, name=f"sweep_D{D}_C{C}_g{g}", d_bound=(-2.0, 3.5, D), channels=C, **gs)
            case("sweep", s, B)
