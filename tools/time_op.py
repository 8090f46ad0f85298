#!/usr/bin/env python
"""Per-kernel CUDA-event timings of the op-level drop-in voxel_pooling (forward / backward) at a given batch."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgv3d_b200 import LiftSplat, get_shape, voxel_pooling, _native as N  # noqa: E402
from sgv3d_b200.synthetic import make_activations, make_mats  # noqa: E402
ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="dair_r50"); ap.add_argument("--batch", type=int, default=4); ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
s = get_shape(a.shape); dev = torch.device("cuda", 0)
mod = LiftSplat(s.x_bound, s.y_bound, s.z_bound, s.d_bound, s.final_dim, s.downsample, s.channels).to(dev)
mats = make_mats(s, a.batch, 1, seed=5, bda="identity")
m = {k: (v.to(dev) if v is not None else None) for k, v in mats.items()}
idx = mod.get_geometry_indices(m["sensor2ego"], m["sensor2virtual"], m["intrin"], m["ida"], m["reference_heights"], m["bda"])
logits, ctx = make_activations(s, a.batch, 1, seed=5, device=dev, generator_device=dev)
feat = (logits.softmax(1).unsqueeze(1) * ctx.unsqueeze(2)).reshape(a.batch, 1, s.channels, s.D, s.fH, s.fW).permute(0, 1, 3, 4, 5, 2).contiguous()
feat.requires_grad_(True)
def fwd():
    with torch.no_grad():
        return voxel_pooling(idx, feat, list(s.grid))
for _ in range(3): fwd()
torch.cuda.synchronize(); N.profile_enable(True); N.profile_report()
for _ in range(a.iters): fwd()
torch.cuda.synchronize(); prof = N.profile_report(); N.profile_enable(False)
tot = sum(t for _, t in prof.values()) / a.iters
ob = s.op_forward_bytes() * a.batch
print(f"[{a.shape} B={a.batch}] op forward: {1e3*tot:.1f} us ({a.batch/(tot*1e-3):.0f} frames/s, {ob/tot/1e6:.0f} GB/s)  " + "  ".join(f"{k}={1e3*t/n:.1f}" for k, (n, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])))
out = voxel_pooling(idx, feat, list(s.grid)); g = torch.randn_like(out)
for _ in range(2): out.backward(g, retain_graph=True)
torch.cuda.synchronize(); N.profile_enable(True); N.profile_report()
for _ in range(a.iters): out.backward(g, retain_graph=True)
torch.cuda.synchronize(); prof = N.profile_report(); N.profile_enable(False)
tot = sum(t for _, t in prof.values()) / a.iters
print(f"[{a.shape} B={a.batch}] op backward: {1e3*tot:.1f} us  " + "  ".join(f"{k}={1e3*t/n:.1f}" for k, (n, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])))
