"""CPU analysis of the voxel-run structure (design aid, not product): runs per pixel, tiles touched per pixel
chunk for row-major chunks vs 2-D pixel blocks, unique pixel rows per reduce tile."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import lift_splat_oracle as O
from sgv3d_b200.shapes import get_shape
from sgv3d_b200.synthetic import make_mats

def voxels(shape, seed, bda="identity"):
    mats = make_mats(shape, 1, 1, seed=seed, bda=bda)
    fr = O.create_frustum(shape.final_dim, shape.downsample, shape.d_bound)
    vs, vc, vn = O.grid_buffers(shape.x_bound, shape.y_bound, shape.z_bound)
    g = O.geometry_matmul(fr, mats["sensor2ego"], mats["sensor2virtual"], mats["intrin"], mats["ida"],
                          mats["reference_heights"], mats.get("bda"))
    idx = O.quantize_np(g.numpy(), vc.numpy(), vs.numpy())[0, 0]      # D,fH,fW,3
    X, Y, Z = shape.grid
    kept = (idx[..., 0] >= 0) & (idx[..., 0] < X) & (idx[..., 1] >= 0) & (idx[..., 1] < Y) & (idx[..., 2] >= 0) & (idx[..., 2] < Z)
    vox = np.where(kept, idx[..., 1] * X + idx[..., 0], -1)            # D,fH,fW
    return vox

def runs_of(vox):
    D, fH, fW = vox.shape
    v = vox.reshape(D, -1)
    start = np.ones_like(v, bool); start[1:] = v[1:] != v[:-1]
    start &= v >= 0
    return v, start      # run starts

def stats(name, seeds=(0, 1, 2), bda="identity", tileshape=(1, 64)):
    shape = get_shape(name)
    X, Y, _ = shape.grid
    fH, fW = shape.fH, shape.fW
    out = []
    for s in seeds:
        vox = voxels(shape, s, bda)
        v, start = runs_of(vox)
        P = fH * fW
        nruns = start.sum(0)                                         # per pixel
        d_idx, p_idx = np.nonzero(start)
        rv = v[d_idx, p_idx]
        ty, tx = tileshape
        tile = (rv // X // ty) * (X // tx) + (rv % X) // tx
        ntiles = (Y // ty) * (X // tx)
        # spans: consecutive runs of a pixel in the same tile
        order = np.lexsort((d_idx, p_idx))
        pp, tt = p_idx[order], tile[order]
        newspan = np.ones(len(pp), bool); newspan[1:] = (pp[1:] != pp[:-1]) | (tt[1:] != tt[:-1])
        spans_per_pixel = newspan.sum() / P
        # unique (tile, pixel) pairs
        pairs = np.unique(tt.astype(np.int64) * P + pp)
        # per tile: entries, unique rows
        ent_per_tile = np.bincount(tile, minlength=ntiles)
        uniq_per_tile = np.bincount((pairs // P).astype(np.int64), minlength=ntiles)
        touched = ent_per_tile > 0
        def chunk_stats(chunk_of_pixel, nch):
            cp = chunk_of_pixel[pp]
            ct = np.unique(cp.astype(np.int64) * ntiles + tt)
            tiles_per_chunk = np.bincount((ct // ntiles).astype(np.int64), minlength=nch)
            cv = np.unique(cp.astype(np.int64) * (X * Y) + rv[order])
            vox_per_chunk = np.bincount((cv // (X * Y)).astype(np.int64), minlength=nch)
            chunks_per_tile = np.bincount((ct % ntiles).astype(np.int64), minlength=ntiles)
            return tiles_per_chunk, vox_per_chunk, chunks_per_tile
        pix = np.arange(P)
        row_major = pix // 128
        h, w = pix // fW, pix % fW
        res = {}
        for nm, (bh, bw) in {"16x8": (16, 8), "8x16": (8, 16), "32x4": (32, 4)}.items():
            nbw = (fW + bw - 1) // bw
            res[nm] = chunk_stats((h // bh) * nbw + (w // bw), ((fH + bh - 1) // bh) * nbw)
        res["row128"] = chunk_stats(row_major, (P + 127) // 128)
        o = dict(seed=s, runs=int(start.sum()), runs_pp=float(nruns.mean()), runs_pp_max=int(nruns.max()), spans_pp=float(spans_per_pixel),
                 pairs=len(pairs), tiles_touched=int(touched.sum()), ent_tile_mean=float(ent_per_tile[touched].mean()),
                 ent_tile_max=int(ent_per_tile.max()), uniq_tile_mean=float(uniq_per_tile[touched].mean()),
                 uniq_tile_p90=float(np.percentile(uniq_per_tile[touched], 90)), uniq_tile_max=int(uniq_per_tile.max()))
        for nm, (tpc, vpc, cpt) in res.items():
            nz = tpc > 0
            o[nm] = dict(chunks=len(tpc), tiles_mean=float(tpc[nz].mean()), tiles_p90=float(np.percentile(tpc[nz], 90)), tiles_max=int(tpc.max()),
                         vox_mean=float(vpc[nz].mean()), vox_p90=float(np.percentile(vpc[nz], 90)), vox_max=int(vpc.max()),
                         chunks_per_tile=float(cpt[touched].mean()))
        out.append(o)
    return out

if __name__ == "__main__":
    import json
    names = sys.argv[1:] or ["dair_r50", "rope3d_r50", "sgv3d_bsm_r50"]
    for nm in names:
        for bda in ("identity", "random"):
            for o in stats(nm, seeds=(0, 1), bda=bda):
                print(nm, bda, json.dumps(o))
