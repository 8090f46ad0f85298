#!/usr/bin/env python
"""Experiment: run the step on nsplit sub-batches concurrently (fork/join streams inside one CUDA graph)."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgv3d_b200 import LiftSplat, get_shape  # noqa: E402
from sgv3d_b200.synthetic import make_activations, make_mats  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="dair_r50"); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--iters", type=int, default=30)
a = ap.parse_args()
s = get_shape(a.shape); dev = torch.device("cuda", 0)
mod = LiftSplat(s.x_bound, s.y_bound, s.z_bound, s.d_bound, s.final_dim, s.downsample, s.channels).to(dev)
sets = []
for i in range(2):
    mats = make_mats(s, a.batch, 1, seed=5 + i, bda="identity")
    md = {"sensor2ego_mats": mats["sensor2ego"].unsqueeze(1).to(dev), "sensor2virtual_mats": mats["sensor2virtual"].unsqueeze(1).to(dev),
          "intrin_mats": mats["intrin"].unsqueeze(1).to(dev), "ida_mats": mats["ida"].unsqueeze(1).to(dev),
          "reference_heights": mats["reference_heights"].unsqueeze(1).to(dev), "bda_mat": mats["bda"].to(dev)}
    logits, ctx = make_activations(s, a.batch, 1, seed=5 + i, device=dev, generator_device=dev)
    sets.append((torch.cat((logits, ctx), 1).contiguous(), md))

def sub(md, lo, hi):
    return {k: (v[lo:hi].contiguous() if v is not None else None) for k, v in md.items()}

for nsplit in (1, 2, 4):
    graphs = []
    for hf, md in sets:
        per = a.batch // nsplit
        parts = [(hf[i * per:(i + 1) * per], sub(md, i * per, (i + 1) * per)) for i in range(nsplit)]
        with torch.no_grad():
            for h, m in parts:
                mod.forward_single_sweep(h, m)
        torch.cuda.synchronize()
        side = [torch.cuda.Stream() for _ in range(nsplit - 1)]
        g = torch.cuda.CUDAGraph()
        outs = []
        with torch.cuda.graph(g), torch.no_grad():
            cur = torch.cuda.current_stream()
            for st in side:
                st.wait_stream(cur)
            for i, (h, m) in enumerate(parts):
                if i == 0:
                    outs.append(mod.forward_single_sweep(h, m))
                else:
                    with torch.cuda.stream(side[i - 1]):
                        outs.append(mod.forward_single_sweep(h, m))
            for st in side:
                cur.wait_stream(st)
        graphs.append((g, outs))
    for _ in range(5):
        for g, _o in graphs: g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters):
        graphs[i % 2][0].replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(f"{a.shape} batch {a.batch} nsplit {nsplit}: {1e3 * ms:.1f} us/step  {a.batch / ms * 1e3:.0f} frames/s", flush=True)
