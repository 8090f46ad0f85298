#!/usr/bin/env python
"""Copy-only microbench: what the platform gives for the end-to-end step WITHOUT any kernel.

Every rank moves, per step, exactly the bytes the end-to-end path of bench.py moves (pinned host -> device for the
head output + matrices, device -> pinned host for the BEV map: 226.5 MB up + 335.5 MB down at 64 DAIR-R50 frames), on
separate upload / download streams with `depth` rotating buffers -- the LiftSplatPipeline structure minus the
CUDA-graph step.  The steady-state step time is max(upload, download) when PCIe is full duplex.  Run it at
N = 1, 2, 4, 8 ranks to separate the platform's host<->device ceiling from pipeline overhead:

    python tools/copy_bench.py                       # 1 rank
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/copy_bench.py

Prints one JSON line per variant (rank 0): default pinned buffers, write-combined upload buffers
(cudaHostAllocWriteCombined), H2D only, D2H only.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgv3d_b200.shapes import get_shape  # noqa: E402


def _wc_pinned(nbytes):
    """cudaHostAlloc(cudaHostAllocWriteCombined) buffer wrapped as a uint8 tensor (kept alive by the caller)."""
    rt = ctypes.CDLL("libcudart.so")
    ptr = ctypes.c_void_p()
    err = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))
    if err != 0 or not ptr.value:
        return None
    arr = (ctypes.c_uint8 * nbytes).from_address(ptr.value)
    return torch.frombuffer(arr, dtype=torch.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="dair_r50")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--depth", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        from sgv3d_b200.sharding import bind_to_gpu_numa_node
        bind_to_gpu_numa_node(local)
        saved = os.dup(1); os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier(device_ids=[local])
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    sh = get_shape(a.shape)
    up_bytes = a.batch * (sh.D + sh.channels) * sh.fH * sh.fW * 4 + a.batch * (4 * 64 + 4 + 64)
    down_bytes = a.batch * sh.channels * sh.grid[0] * sh.grid[1] * 4

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def run(variant, do_up, do_down, wc):
        d_in = [torch.empty(up_bytes, dtype=torch.uint8, device=dev) for _ in range(a.depth)]
        d_out = [torch.empty(down_bytes, dtype=torch.uint8, device=dev) for _ in range(a.depth)]
        h_in = None
        if wc:
            h_in = _wc_pinned(up_bytes)
        if h_in is None:
            if wc:
                return None
            h_in = torch.empty(up_bytes, dtype=torch.uint8).pin_memory()
        h_in.fill_(1)
        h_out = [torch.empty(down_bytes, dtype=torch.uint8).pin_memory() for _ in range(a.depth)]
        up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        done = [torch.cuda.Event() for _ in range(a.depth)]

        def loop(n):
            for i in range(n):
                k = i % a.depth
                if do_up:
                    with torch.cuda.stream(up):
                        up.wait_event(done[k])
                        d_in[k].copy_(h_in, non_blocking=True)
                if do_down:
                    with torch.cuda.stream(down):
                        h_out[k].copy_(d_out[k], non_blocking=True)
                        done[k].record(down)
            torch.cuda.synchronize()
        loop(4)
        barrier()
        t0 = time.perf_counter()
        loop(a.steps)
        ms = 1e3 * (time.perf_counter() - t0) / a.steps
        barrier()
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        if rank == 0:
            gb = ((up_bytes if do_up else 0) + (down_bytes if do_down else 0)) / 1e9
            print(json.dumps({"copy_bench": variant, "n_gpus": world, "shape": sh.name, "frames_per_step_per_gpu": a.batch,
                              "h2d_bytes_per_step": up_bytes if do_up else 0, "d2h_bytes_per_step": down_bytes if do_down else 0,
                              "ms_per_step_max_over_ranks": ms, "frames_per_s_aggregate": world * a.batch / (ms * 1e-3),
                              "GBs_per_rank": gb / (ms * 1e-3), "GBs_aggregate": world * gb / (ms * 1e-3)}), flush=True)

    run("h2d+d2h pinned", True, True, False)
    run("h2d+d2h write-combined upload buffer", True, True, True)
    run("h2d only", True, False, False)
    run("d2h only", False, True, False)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
