"""CPU analysis for the block-centric design: footprint (distinct voxels) of 8x8 pixel half-blocks, strips of 32 voxels
touched, blocks contributing per voxel."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tools.analysis.plan_stats import voxels, runs_of
from sgv3d_b200.shapes import get_shape

def stats(name, seed, bda, bh=8, bw=8):
    shape = get_shape(name)
    X, Y, _ = shape.grid
    fH, fW = shape.fH, shape.fW
    vox = voxels(shape, seed, bda)
    v, start = runs_of(vox)
    P = fH * fW
    d_idx, p_idx = np.nonzero(start)
    rv = v[d_idx, p_idx]
    h, w = p_idx // fW, p_idx % fW
    nbw = (fW + bw - 1) // bw
    nb = ((fH + bh - 1) // bh) * nbw
    blk = (h // bh) * nbw + (w // bw)
    V = X * Y
    bv = np.unique(blk.astype(np.int64) * V + rv)
    slots = np.bincount((bv // V).astype(np.int64), minlength=nb)
    bs = np.unique(blk.astype(np.int64) * (V // 32) + rv // 32)
    strips = np.bincount((bs // (V // 32)).astype(np.int64), minlength=nb)
    runs_b = np.bincount(blk, minlength=nb)
    fan = np.bincount((bv % V).astype(np.int64), minlength=V)
    nz = slots > 0
    return dict(shape=name, bda=bda, seed=seed, blocks=int(nb), live_blocks=int(nz.sum()), runs=len(rv),
                runs_blk_mean=float(runs_b[nz].mean()), runs_blk_max=int(runs_b.max()),
                slots_mean=float(slots[nz].mean()), slots_p90=float(np.percentile(slots[nz], 90)), slots_p99=float(np.percentile(slots[nz], 99)), slots_max=int(slots.max()),
                strips_mean=float(strips[nz].mean()), strips_max=int(strips.max()),
                partial_rows=int(slots.sum()), voxels_hit=int((fan > 0).sum()), fan_mean=float(fan[fan > 0].mean()), fan_max=int(fan.max()))

if __name__ == "__main__":
    names = sys.argv[1:] or ["dair_r50", "rope3d_r50", "sgv3d_bsm_r50", "rope3d_r101_256", "rope3d_r101_140", "sgv3d_bsm_r101", "rope3d_native"]
    for nm in names:
        for bda in ("identity", "random"):
            for s in (0, 1, 2):
                print(json.dumps(stats(nm, s, bda)))
