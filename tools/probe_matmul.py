#!/usr/bin/env python
"""Which rounding order does torch's CUDA matmul use for a batch of 4x4 @ 4x4 fp32 products, by batch count?
Candidates are emulated in float64 (a*b is exact there; one rounding to fp32 per emulated operation)."""
import itertools, torch
torch.manual_seed(0)
dev = "cuda"
f32, f64 = torch.float32, torch.float64

def rn(x):
    return x.to(f32).to(f64)

def cand(A, B):
    a, b = A.to(f64), B.to(f64)
    p = [a[..., :, k:k + 1] * b[..., k:k + 1, :] for k in range(4)]   # exact products (..., 4, 4)
    out = {}
    out["FMA_asc"] = rn(p[3] + rn(p[2] + rn(p[1] + rn(p[0]))))
    out["FMA_desc"] = rn(p[0] + rn(p[1] + rn(p[2] + rn(p[3]))))
    out["SEQ"] = rn(rn(rn(rn(p[0]) + rn(p[1])) + rn(p[2])) + rn(p[3]))
    out["PAIR"] = rn(rn(p[1] + rn(p[0])) + rn(p[3] + rn(p[2])))
    out["PAIR2"] = rn(rn(p[2] + rn(p[0])) + rn(p[3] + rn(p[1])))
    return {k: v.to(f32) for k, v in out.items()}

for shape in [(1, 1), (2, 1), (4, 1), (8, 1), (16, 1), (32, 1), (64, 1), (128, 1), (256, 1), (8, 6), (64, 2)]:
    A = torch.randn(*shape, 4, 4, device=dev); B = torch.randn(*shape, 4, 4, device=dev)
    A[..., 3, :] = torch.tensor([0., 0., 0., 1.], device=dev)
    want = A.matmul(B)
    res = {k: int((v.view(torch.int32) != want.view(torch.int32)).sum()) for k, v in cand(A, B).items()}
    print(shape, "mismatching floats of", want.numel(), res, flush=True)
