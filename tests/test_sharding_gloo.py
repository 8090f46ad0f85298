"""N > 1 host logic on CPU: world_size-2 gloo run of the frame sharding ("replicas only").

Each rank takes its contiguous slice of a global batch, runs the (CPU) oracle of the path on it, and the
gathered per-rank BEV maps must equal the unsharded run bit for bit -- the path has no cross-frame term, so
sharding may not change a single bit.  The timing / counter reductions bench.py uses are checked too."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lift_splat_oracle as O
from sgv3d_b200.shapes import get_shape
from sgv3d_b200.sharding import max_over_ranks, shard_bounds, shard_frames, shard_mats, sum_over_ranks
from sgv3d_b200.synthetic import make_activations, make_mats

GLOBAL_BATCH = 5   # odd on purpose: ranks get 3 and 2 frames
NUM_CAMS = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs():
    shape = get_shape("tiny")
    mats = make_mats(shape, GLOBAL_BATCH, NUM_CAMS, seed=7, bda="random")
    logits, ctx = make_activations(shape, GLOBAL_BATCH, NUM_CAMS, seed=7)
    md = {"sensor2ego_mats": mats["sensor2ego"], "sensor2virtual_mats": mats["sensor2virtual"],
          "intrin_mats": mats["intrin"], "ida_mats": mats["ida"], "reference_heights": mats["reference_heights"],
          "bda_mat": mats["bda"]}
    return shape, md, logits, ctx


def _oracle_bev(shape, md, logits, ctx):
    fr = O.create_frustum(shape.final_dim, shape.downsample, shape.d_bound)
    vs, vc, vn = O.grid_buffers(shape.x_bound, shape.y_bound, shape.z_bound)
    mats = {"sensor2ego": md["sensor2ego_mats"], "sensor2virtual": md["sensor2virtual_mats"], "intrin": md["intrin_mats"],
            "ida": md["ida_mats"], "reference_heights": md["reference_heights"], "bda": md["bda_mat"]}
    bev, _, _ = O.lift_splat_forward(logits, ctx, fr, mats, vc, vs, vn)
    return bev.numpy()


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shape, md, logits, ctx = _inputs()
        lo, hi = shard_bounds(GLOBAL_BATCH, world, rank)
        bev = _oracle_bev(shape, shard_mats(md, lo, hi), shard_frames(logits, NUM_CAMS, lo, hi),
                          shard_frames(ctx, NUM_CAMS, lo, hi))
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, bev))
        ms = max_over_ranks([10.0 + rank, 5.0 - rank])
        frames = sum_over_ranks([hi - lo])
        if rank == 0:
            np.savez(os.path.join(out_dir, "r0.npz"), bev=np.concatenate([g[2] for g in gathered], 0),
                     bounds=np.array([[g[0], g[1]] for g in gathered]), ms=np.array(ms), frames=np.array(frames))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_partition():
    for n in (0, 1, 5, 8, 64, 67):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_reductions_without_process_group_are_identity():
    assert max_over_ranks([1.5, 2.0]) == [1.5, 2.0]
    assert sum_over_ranks([3]) == [3.0]


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharding_matches_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "r0.npz")
    shape, md, logits, ctx = _inputs()
    want = _oracle_bev(shape, md, logits, ctx)
    assert got["bounds"].tolist() == [[0, 3], [3, 5]]
    assert got["bev"].shape == want.shape
    assert np.array_equal(got["bev"], want), "sharding changed the result"
    assert got["ms"].tolist() == [11.0, 5.0]          # max over ranks of (10 + r, 5 - r)
    assert got["frames"].tolist() == [float(GLOBAL_BATCH)]
