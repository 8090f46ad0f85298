#!/usr/bin/env python
"""Test infrastructure (not collected by pytest): times the reference's own voxel_pooling kernel
(ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu, compiled unmodified into oracle/_ref by oracle/Makefile)
next to this library's op-level drop-in on the same materialised inputs, on the GPU.
    python tests/bench_reference_kernel.py [--shape dair_r50] [--frames 4] > profiles/reference_kernel_r01.json"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle as CO  # noqa: E402
from sgv3d_b200 import LiftSplat, get_shape, voxel_pooling  # noqa: E402
from sgv3d_b200.synthetic import make_activations, make_mats  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="dair_r50"); ap.add_argument("--frames", type=int, default=4)
a = ap.parse_args()
s = get_shape(a.shape); dev = torch.device("cuda", 0); nb = a.frames
assert CO.reference_kernel_available(), "build oracle/_ref first (make -C oracle ref)"
mod = LiftSplat(s.x_bound, s.y_bound, s.z_bound, s.d_bound, s.final_dim, s.downsample, s.channels).to(dev)
mats = make_mats(s, nb, 1, seed=5, bda="identity")
dm = {k: v.to(dev) for k, v in mats.items()}
logits, ctx = make_activations(s, nb, 1, seed=5, device=dev, generator_device=dev)
idx = mod.get_geometry_indices(dm["sensor2ego"], dm["sensor2virtual"], dm["intrin"], dm["ida"], dm["reference_heights"], dm["bda"])
D, C = s.D, s.channels
feat = (logits.softmax(1).unsqueeze(1) * ctx.unsqueeze(2)).reshape(nb, 1, C, D, s.fH, s.fW).permute(0, 1, 3, 4, 5, 2).contiguous()
X, Y, Z = s.grid
npts = s.points_per_frame
o = torch.zeros(nb, Y, X, C, device=dev)
pm = torch.full((nb, npts, 3), -1, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def ref():
    o.zero_(); pm.fill_(-1)      # what VoxelPooling.forward does before the launch (voxel_pooling.py:37-40)
    CO.reference_voxel_pooling_forward(nb, npts, C, X, Y, Z, idx.data_ptr(), feat.data_ptr(), o.data_ptr(), pm.data_ptr(), stream)


def ours():
    with torch.no_grad():
        return voxel_pooling(idx, feat, list(s.grid))


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ms_ref, ms_ours = timeit(ref), timeit(ours)
got = ours().permute(0, 2, 3, 1)
ref()
torch.cuda.synchronize()
ob = s.op_forward_bytes() * nb
print(json.dumps({"shape": s.name, "frames": nb, "op_level_bytes": ob,
                  "reference_kernel_sm100a": {"ms": ms_ref, "frames_per_s": nb / ms_ref * 1e3, "GBs": ob / ms_ref / 1e6,
                                              "note": "atomicAdd kernel recompiled unmodified, incl. its output / pos_memo fills"},
                  "sgv3d_b200_voxel_pooling": {"ms": ms_ours, "frames_per_s": nb / ms_ours * 1e3, "GBs": ob / ms_ours / 1e6},
                  "max_abs_diff": float((got - o).abs().max()), "max_abs_ref": float(o.abs().max())}))
