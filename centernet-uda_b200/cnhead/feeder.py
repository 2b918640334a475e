"""Pinned-host -> HBM staging for the head path: the role ``batch[k] = batch[k].to(device)`` plays in the
reference's training loop (train.py:148-150), made asynchronous.

``HostFeeder`` owns a copy stream and ``depth`` device slots.  ``put()`` enqueues the H2D copies of one
step's pinned host tensors into the next free slot on the copy stream; ``get()`` makes the caller's
stream wait for the oldest staged slot and hands out its device tensors.  With ``depth >= 2`` the copy
of step i+1 runs on the copy engine while the kernels of step i run on the SMs, so a step costs
max(copy, compute) instead of their sum.  Nothing here touches the data: it is plumbing around
cudaMemcpyAsync (torch's ``copy_(non_blocking=True)``), not part of the kernels.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch


class HostFeeder:
    def __init__(self, device, depth: int = 2):
        if depth < 1:
            raise ValueError("HostFeeder: depth must be >= 1")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostFeeder: a CUDA device is required (there is no CPU path)")
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots: List[Tuple[Dict[str, torch.Tensor], ...]] = [None] * depth   # device tensors per slot
        self._ready = [torch.cuda.Event() for _ in range(depth)]                  # copies of the slot landed
        self._free = [None] * depth                                               # consumer finished with the slot
        self._head = 0                                                            # next slot to fill
        self._tail = 0                                                            # next slot to hand out
        self.bytes_per_put = 0

    def _device_like(self, dicts):
        return tuple({k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in d.items()} for d in dicts)

    def put(self, *dicts: Dict[str, torch.Tensor]) -> None:
        """Stage one step: every tensor must live in pinned host memory (checked once per slot)."""
        if self._head - self._tail >= self.depth:
            raise RuntimeError("HostFeeder: all slots are staged; call get() first")
        i = self._head % self.depth
        if self._slots[i] is None or any(dv.shape != hv.shape for dd, hd in zip(self._slots[i], dicts)
                                         for dv, hv in zip(dd.values(), hd.values())):
            for d in dicts:
                for k, v in d.items():
                    if not v.is_pinned():
                        raise RuntimeError(f"HostFeeder: host tensor '{k}' is not pinned")
            self._slots[i] = self._device_like(dicts)
        n = 0
        with torch.cuda.stream(self.copy_stream):
            if self._free[i] is not None:
                self.copy_stream.wait_event(self._free[i])       # the consumer of the slot's previous content is done
            for dev_d, host_d in zip(self._slots[i], dicts):
                for k, v in host_d.items():
                    dev_d[k].copy_(v, non_blocking=True)
                    n += v.numel() * v.element_size()
            self._ready[i].record(self.copy_stream)
        self.bytes_per_put = n
        self._head += 1

    def get(self) -> Tuple[Dict[str, torch.Tensor], ...]:
        """Device tensors of the oldest staged step; the current stream waits for their copies."""
        if self._tail >= self._head:
            raise RuntimeError("HostFeeder: nothing staged; call put() first")
        i = self._tail % self.depth
        torch.cuda.current_stream(self.device).wait_event(self._ready[i])
        self._tail += 1
        return self._slots[i]

    def release(self) -> None:
        """Call after the last kernel that reads the most recently handed-out slot has been enqueued."""
        i = (self._tail - 1) % self.depth
        ev = self._free[i] or torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[i] = ev
