#!/usr/bin/env python
"""Dump (A, torch.linalg.inv_ex(A) on CUDA) pairs for 4x4 calibration-like and random matrices, and check that the
result does not depend on the batch size.  Output: gpurun_out/inverse_probe.npz (analysed offline)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgv3d_b200 import get_shape
from sgv3d_b200.synthetic import make_mats
from sgv3d_b200.view_transform import _inverse
torch.manual_seed(0)
mats = []
for fam in ("dair_r50", "rope3d_r50"):
    m = make_mats(get_shape(fam), 512, 1, seed=3, bda="random")
    for k in ("ida", "intrin", "sensor2virtual", "sensor2ego"):
        mats.append(m[k].reshape(-1, 4, 4))
mats.append(torch.randn(4096, 4, 4))
mats.append(torch.randn(4096, 4, 4) * torch.logspace(-2, 3, 4).view(1, 1, 4))
A = torch.cat(mats, 0).float().cuda().contiguous()
inv_all = _inverse(A)
bad = {}
for bs in (1, 3, 8, 96, 1000):
    n = 0
    for i in range(0, 2000, bs):
        blk = A[i:i + bs]
        n += int((_inverse(blk).view(torch.int32) != inv_all[i:i + bs].view(torch.int32)).sum())
    bad[bs] = n
print("mismatching floats vs one big batch, by batch size:", bad)
lu, piv = torch.linalg.lu_factor(A)
np.savez_compressed("gpurun_out/inverse_probe.npz", A=A.cpu().numpy(), inv=inv_all.cpu().numpy(), lu=lu.cpu().numpy(), piv=piv.cpu().numpy())
print("saved", A.shape)
