"""Host-to-host serving loop around the fused lift-splat: pinned host inputs -> H2D -> CUDA-graph replay of
``LiftSplat.forward_single_sweep`` (lss_fpn.py:462-495) -> D2H of the BEV map into pinned host memory.

Three streams (upload, compute, download) and ``depth`` rotating slots, so that the upload of step i+1 and
the download of step i-1 overlap the kernels of step i; PCIe is full duplex, so the steady-state step time
is max(upload, compute, download) instead of their sum.  Every byte still crosses the bus on every step.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .view_transform import LiftSplat, LiftSplatGraph

__all__ = ["LiftSplatPipeline"]

_MAT_KEYS = ("sensor2ego_mats", "sensor2virtual_mats", "intrin_mats", "ida_mats", "reference_heights", "bda_mat")


class _Slot:
    def __init__(self, module: LiftSplat, example_hf: torch.Tensor, example_mats: Dict[str, Optional[torch.Tensor]],
                 device, sweep_index: int):
        self.hf = torch.empty(example_hf.shape, dtype=torch.float32, device=device)
        self.mats = {k: (torch.empty(v.shape, dtype=torch.float32, device=device) if v is not None else None)
                     for k, v in example_mats.items()}
        # capture needs valid matrices (inverse of garbage is harmless but NaNs slow nothing down; still, be tidy)
        self.hf.copy_(example_hf)
        for k, v in example_mats.items():
            if v is not None:
                self.mats[k].copy_(v)
        self.graph = LiftSplatGraph(module, self.hf, self.mats, sweep_index)
        self.bev_host = torch.empty(self.graph.bev.shape, dtype=torch.float32).pin_memory()
        self.uploaded = torch.cuda.Event()
        self.computed = torch.cuda.Event()
        self.downloaded = torch.cuda.Event()
        self.busy = False


class LiftSplatPipeline:
    """``submit(height_feature_host, mats_dict_host)`` queues one batch and returns its slot;
    ``result(slot)`` waits for that batch and returns the pinned host BEV map (valid until the slot is
    submitted again, i.e. for ``depth - 1`` further submits)."""

    def __init__(self, module: LiftSplat, example_height_feature: torch.Tensor,
                 example_mats: Dict[str, Optional[torch.Tensor]], depth: int = 3, sweep_index: int = 0,
                 device: Optional[torch.device] = None):
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if device.type != "cuda":
            raise RuntimeError("sgv3d_b200 runs on CUDA devices only (no CPU fallback)")
        self.device = device
        self.depth = depth
        with torch.cuda.device(device):
            self.up = torch.cuda.Stream(device=device)
            self.compute = torch.cuda.Stream(device=device)
            self.down = torch.cuda.Stream(device=device)
            self.slots: List[_Slot] = [_Slot(module, example_height_feature, example_mats, device, sweep_index)
                                       for _ in range(depth)]
        self._next = 0
        hf = self.slots[0].hf
        self.h2d_bytes = hf.numel() * 4 + sum(v.numel() * 4 for v in self.slots[0].mats.values() if v is not None)
        self.d2h_bytes = self.slots[0].bev_host.numel() * 4
        torch.cuda.synchronize(device)

    def submit(self, height_feature_host: torch.Tensor, mats_host: Dict[str, Optional[torch.Tensor]]) -> int:
        i = self._next
        self._next = (i + 1) % self.depth
        s = self.slots[i]
        if s.busy:
            # the slot's previous download must be over before its buffers are reused; its previous
            # compute precedes that download, so one wait covers both the input and the output buffer
            self.up.wait_event(s.downloaded)
        with torch.cuda.stream(self.up):
            s.hf.copy_(height_feature_host, non_blocking=True)
            for k in _MAT_KEYS:
                v = mats_host.get(k)
                if v is not None and s.mats.get(k) is not None:
                    s.mats[k].copy_(v, non_blocking=True)
            s.uploaded.record(self.up)
        self.compute.wait_event(s.uploaded)
        with torch.cuda.stream(self.compute):
            s.graph()
            s.computed.record(self.compute)
        self.down.wait_event(s.computed)
        with torch.cuda.stream(self.down):
            s.bev_host.copy_(s.graph.bev, non_blocking=True)
            s.downloaded.record(self.down)
        s.busy = True
        return i

    def result(self, slot: int) -> torch.Tensor:
        s = self.slots[slot]
        s.downloaded.synchronize()
        return s.bev_host

    def drain(self) -> None:
        for s in self.slots:
            if s.busy:
                s.downloaded.synchronize()
