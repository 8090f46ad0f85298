"""Frame sharding for one-process-per-GPU runs ("replicas only").

No term of the view transform couples two frames (the reference kernel indexes by
``batch_idx = pt_idx / num_points``, ops/voxel_pooling/src/voxel_pooling_forward_cuda.cu:19, and
lss_fpn.py:462-495 has no batch-crossing op), so a global batch is cut into contiguous per-rank
slices and each rank runs the whole path on its slice.  There is no collective on the data path;
``torch.distributed`` is touched only to agree on timings / counters (``max_over_ranks``,
``sum_over_ranks``) -- NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

__all__ = ["shard_bounds", "shard_mats", "shard_frames", "max_over_ranks", "sum_over_ranks", "bind_to_gpu_numa_node"]

# keys of the reference's ``mats_dict`` (dataset/nusc_mv_det_dataset.py:864-871) and their frame axis
_PER_FRAME_KEYS = ("sensor2ego_mats", "sensor2virtual_mats", "intrin_mats", "ida_mats", "reference_heights",
                   "bda_mat")


def shard_bounds(num_frames: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the frames rank ``rank`` owns: contiguous, balanced to within one frame, the
    first ``num_frames % world_size`` ranks take the extra frame."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    if num_frames < 0:
        raise ValueError("num_frames < 0")
    base, extra = divmod(num_frames, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_mats(mats_dict: Dict[str, Optional[torch.Tensor]], lo: int, hi: int) -> Dict[str, Optional[torch.Tensor]]:
    """Slice every per-frame entry of a reference-style ``mats_dict`` to frames [lo, hi)."""
    out = {}
    for k, v in mats_dict.items():
        out[k] = v[lo:hi] if (v is not None and k in _PER_FRAME_KEYS) else v
    return out


def shard_frames(per_camera: torch.Tensor, num_cams: int, lo: int, hi: int) -> torch.Tensor:
    """Slice a ``(B*Nc, ...)`` per-camera tensor (height-net output, context, ...) to frames [lo, hi)."""
    return per_camera[lo * num_cams:hi * num_cams]


def _reduce(values, op, group=None, device=None) -> list:
    import torch.distributed as dist
    vals = [float(v) for v in values]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return vals
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
            else torch.device("cpu")
    t = torch.tensor(vals, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op, group=group)
    return [float(x) for x in t.tolist()]


def max_over_ranks(values, group=None, device=None) -> list:
    """Element-wise max over ranks of a short list of scalars (per-rank device timings)."""
    import torch.distributed as dist
    return _reduce(values, dist.ReduceOp.MAX, group, device)


def sum_over_ranks(values, group=None, device=None) -> list:
    """Element-wise sum over ranks (frames processed, kernels launched)."""
    import torch.distributed as dist
    return _reduce(values, dist.ReduceOp.SUM, group, device)


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """One process per GPU: pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned
    host staging buffers (first touch) and the threads that fill them are local to the GPU's PCIe root.  Without
    it every rank's buffers may land on one socket and the host<->device copies of an 8-GPU box contend for
    the inter-socket link.  Best effort: returns the node, or None when the topology cannot be read (then
    nothing is changed).  Call it before allocating pinned memory."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError):
        return None
