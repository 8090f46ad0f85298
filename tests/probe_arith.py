#!/usr/bin/env python
"""GPU probe: in which order does torch's CUDA ``matmul`` (cuBLAS bmm) evaluate the 4-term dot
products of the reference's geometry (lss_fpn.py:361-362,367-369,392,398)?

Runs the oracle's torch port of get_geometry on the GPU stage by stage, then replays every stage on
the CPU from the GPU's own stage inputs under several candidate orders (fp64-emulated FMA) and counts
bit mismatches.  Also compares our kernel (SEQ / FMA) end to end.  Writes gpurun_out/arith_probe.json.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lift_splat_oracle as O  # noqa: E402
from sgv3d_b200 import get_shape  # noqa: E402
from sgv3d_b200.synthetic import make_mats  # noqa: E402

f32, f64 = np.float32, np.float64


def fma(a, b, c):
    return (a.astype(f64) * b.astype(f64) + c.astype(f64)).astype(f32)


def candidates(m, v):
    """m: (..., 4) row, v: list of 4 arrays; returns dict name -> result."""
    a = [m[..., k] for k in range(4)]
    p = [(a[k] * v[k]).astype(f32) for k in range(4)]
    out = {}
    out["seq"] = ((p[0] + p[1]) + p[2]) + p[3]
    out["fma_fwd"] = fma(a[3], v[3], fma(a[2], v[2], fma(a[1], v[1], p[0])))
    out["fma_rev"] = fma(a[0], v[0], fma(a[1], v[1], fma(a[2], v[2], p[3])))
    out["pair_seq"] = (p[0] + p[1]) + (p[2] + p[3])
    out["pair_fma"] = fma(a[1], v[1], p[0]) + fma(a[3], v[3], p[2])
    out["seq_rev"] = ((p[3] + p[2]) + p[1]) + p[0]
    out["fma_fwd_f64acc"] = (a[0].astype(f64) * v[0] + a[1].astype(f64) * v[1] + a[2].astype(f64) * v[2]
                             + a[3].astype(f64) * v[3]).astype(f32)
    return out


def stage_report(name, mat, vec, got):
    """mat (B,Nc,4,4) np, vec (B,Nc,D,H,W,4) np, got (B,Nc,D,H,W,4) np"""
    res = {}
    B, Nc = mat.shape[:2]
    v = [vec[..., k] for k in range(4)]
    for r in range(4):
        row = mat[:, :, r].reshape(B, Nc, 1, 1, 1, 4)
        with np.errstate(all="ignore"):
            for cname, val in candidates(row, v).items():
                val = np.broadcast_to(val, got[..., r].shape)
                bad = int((val.view(np.int32) != np.ascontiguousarray(got[..., r]).view(np.int32)).sum())
                res[cname] = res.get(cname, 0) + bad
    res["n"] = int(got.size)
    print(name, res, flush=True)
    return res


def main():
    out = {"torch": torch.__version__, "device": torch.cuda.get_device_name(0)}
    for shape_name in ("dair_r50", "rope3d_r50", "small"):
        s = get_shape(shape_name)
        B = 2
        mats = make_mats(s, B, 1, seed=77, bda="random")
        fr = O.create_frustum(s.final_dim, s.downsample, s.d_bound)
        dev = {k: v.cuda() for k, v in mats.items()}
        frd = fr.cuda()
        D, fh, fw = fr.shape[:3]
        # stage-by-stage on the GPU (same calls as geometry_matmul)
        ida_inv, mv, me = O.camera_matrices(dev["sensor2ego"], dev["sensor2virtual"], dev["intrin"], dev["ida"])
        ida_inv_c, mv_c, me_c = O.camera_matrices(mats["sensor2ego"], mats["sensor2virtual"], mats["intrin"], mats["ida"])
        rep = {"prep_bits_differ_cpu_vs_gpu": {
            "ida_inv": int((ida_inv.cpu() != ida_inv_c).sum()), "m_virtual": int((mv.cpu() != mv_c).sum()),
            "m_ego": int((me.cpu() != me_c).sum())}}
        p0 = ida_inv.view(B, 1, 1, 1, 1, 4, 4).matmul(frd.unsqueeze(-1))
        rh = dev["reference_heights"].view(B, 1, 1, 1, 1, 1).expand(B, 1, D, fh, fw, 1)
        hgt = -1 * p0[..., 2, :] + rh
        ray = p0.clone()
        ray[..., 2, :] = 10
        ray = torch.cat((ray[..., :2, :] * ray[..., 2:3, :], ray[..., 2:, :]), dim=-2)
        pv = mv.view(B, 1, 1, 1, 1, 4, 4).matmul(ray)
        ratio = hgt[..., 0] / pv[..., 1, 0]
        pe = pv * ratio.view(B, 1, D, fh, fw, 1, 1)
        pe[..., 3, :] = 1
        pg = me.view(B, 1, 1, 1, 1, 4, 4).matmul(pe)
        pb = dev["bda"].view(B, 1, 1, 1, 1, 4, 4).expand(B, 1, 1, 1, 1, 4, 4) @ pg
        n = lambda t: t.squeeze(-1).cpu().numpy()
        frn = np.broadcast_to(fr.numpy()[None, None], (B, 1, D, fh, fw, 4))
        rep["ida"] = stage_report(f"{shape_name} stage ida", ida_inv.cpu().numpy(), frn, n(p0))
        rep["virtual"] = stage_report(f"{shape_name} stage virtual", mv.cpu().numpy(), n(ray), n(pv))
        rep["ego"] = stage_report(f"{shape_name} stage ego", me.cpu().numpy(), n(pe), n(pg))
        rep["bda"] = stage_report(f"{shape_name} stage bda", np.broadcast_to(mats["bda"].numpy()[:, None], (B, 1, 4, 4)), n(pg), n(pb))
        # elementwise stages: exact IEEE?
        with np.errstate(all="ignore"):
            rep["ratio_ieee_mismatch"] = int(((n(hgt.unsqueeze(-1))[..., 0] / n(pv)[..., 1]).astype(f32).view(np.int32)
                                              != ratio.cpu().numpy().view(np.int32)).sum())
        # end-to-end: our kernel vs the device oracle
        vs, vc, vn = O.grid_buffers(s.x_bound, s.y_bound, s.z_bound)
        geom_dev = O.geometry_matmul(frd, dev["sensor2ego"], dev["sensor2virtual"], dev["intrin"], dev["ida"],
                                     dev["reference_heights"], dev["bda"])
        idx_dev = O.quantize(geom_dev, vc.cuda(), vs.cuda()).cpu().numpy()
        geom_cpu = O.geometry_matmul(fr, mats["sensor2ego"], mats["sensor2virtual"], mats["intrin"], mats["ida"],
                                     mats["reference_heights"], mats["bda"])
        idx_cpu = O.quantize(geom_cpu, vc, vs).numpy()
        rep["points"] = int(idx_dev.size // 3)
        rep["torch_cuda_vs_torch_cpu_idx_points_differ"] = int((idx_dev != idx_cpu).any(-1).sum())
        from sgv3d_b200.view_transform import geometry_indices
        for arith, nm in ((0, "SEQ"), (1, "FMA"), (2, "PAIR")):
            idx_k, xyz_k = geometry_indices(frd, dev["sensor2ego"], dev["sensor2virtual"], dev["intrin"], dev["ida"],
                                            dev["reference_heights"], dev["bda"], vc, vs, arith=arith, return_xyz=True)
            rep[f"kernel_{nm}_vs_torch_cuda_idx_points_differ"] = int((idx_k.cpu().numpy() != idx_dev).any(-1).sum())
            rep[f"kernel_{nm}_vs_torch_cuda_xyz_floats_differ"] = int(
                (xyz_k.cpu().numpy().view(np.int32) != geom_dev.contiguous().cpu().numpy().view(np.int32)).sum())
        print(shape_name, json.dumps(rep), flush=True)
        out[shape_name] = rep
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/arith_probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
