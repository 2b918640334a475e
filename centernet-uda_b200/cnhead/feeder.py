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


class PinnedArena:
    """Pinned host tensors carved from ONE large page-locked allocation.

    Measured on the B200 boxes (tools/h2d_probe.py): the step's eight tensors copy at 21-42 GB/s from
    separate ``pin_memory()`` allocations (erratic between boxes and runs) and at a steady 50.5 GB/s from
    slices of a single 256 MiB pinned buffer -- large page-locked allocations are backed by large,
    physically contiguous pages, which the copy engine and the IOMMU handle at full PCIe rate."""

    MIN_BYTES = 256 << 20
    ALIGN = 4096

    def __init__(self, nbytes: int = 0):
        self.buf = torch.empty(max(int(nbytes), self.MIN_BYTES), dtype=torch.uint8).pin_memory()
        self.off = 0

    @classmethod
    def bytes_for(cls, tensors: Sequence[torch.Tensor]) -> int:
        return sum((t.numel() * t.element_size() + cls.ALIGN - 1) // cls.ALIGN * cls.ALIGN for t in tensors)

    def take_like(self, t: torch.Tensor, copy: bool = True) -> torch.Tensor:
        n = t.numel() * t.element_size()
        if self.off + n > self.buf.numel():
            raise RuntimeError("PinnedArena: out of space")
        out = self.buf[self.off:self.off + n].view(t.dtype).view(t.shape)
        self.off += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if copy:
            out.copy_(t)
        return out


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
        self._handed = [False] * depth                                            # get() handed it out, release() not yet called
        self._head = 0                                                            # next slot to fill
        self._tail = 0                                                            # next slot to hand out
        self.bytes_per_put = 0
        self._plans = {}                                                          # (slot, host set) -> copy plan

    @staticmethod
    def pinned_sets(sets: Sequence[Sequence[Dict[str, torch.Tensor]]]) -> List[Tuple[Dict[str, torch.Tensor], ...]]:
        """Host staging for ``put()``: the given sets (each a sequence of dicts of CPU tensors) re-created, with
        their contents, as slices of one PinnedArena.  A data loader would write its batches straight into
        such buffers (``take_like(..., copy=False)``)."""
        flat = [t for st in sets for d in st for t in d.values()]
        arena = PinnedArena(PinnedArena.bytes_for(flat))
        return [tuple({k: arena.take_like(v) for k, v in d.items()} for d in st) for st in sets]

    def _device_like(self, dicts):
        return tuple({k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in d.items()} for d in dicts)

    @staticmethod
    def _span(dicts):
        """(first byte, byte span, relative layout) when every tensor is a contiguous slice of ONE host storage
        packed densely enough to be shipped as a single copy (PinnedArena sets are); else None."""
        ts = [v for d in dicts for v in d.values()]
        base = ts[0].untyped_storage().data_ptr()
        if any(v.untyped_storage().data_ptr() != base or not v.is_contiguous() for v in ts):
            return None
        offs = [v.data_ptr() - base for v in ts]
        sizes = [v.numel() * v.element_size() for v in ts]
        lo = min(offs)
        span = max(o + n for o, n in zip(offs, sizes)) - lo
        if lo % 16 or any((o - lo) % 16 for o in offs) or span > 1.25 * sum(sizes) + (1 << 16):
            return None
        return lo, span, tuple((o - lo, v.dtype, tuple(v.shape)) for o, v in zip(offs, ts))

    def put(self, *dicts: Dict[str, torch.Tensor]) -> None:
        """Stage one step: every tensor must live in pinned host memory (checked once per slot).  A set that
        is one dense span of a single pinned storage travels as ONE copy into a mirrored device span."""
        if self._head - self._tail >= self.depth:
            raise RuntimeError("HostFeeder: all slots are staged; call get() first")
        i = self._head % self.depth
        if self._handed[i]:
            # overwriting a slot whose consumer never told us it was done would race the copy against its kernels
            raise RuntimeError("HostFeeder: the slot to be refilled was handed out by get() but never release()d")
        key = tuple(id(v) for d in dicts for v in d.values())
        plan = self._plans.get((i, key))
        if plan is None:                                     # first time this host set meets this slot
            for d in dicts:
                for k, v in d.items():
                    if not v.is_pinned():
                        raise RuntimeError(f"HostFeeder: host tensor '{k}' is not pinned")
            sp = self._span(dicts)
            slot = self._slots[i]
            if sp is not None:
                lo, span, layout = sp
                if slot is None or slot[2] != layout:
                    arena = torch.empty(span, dtype=torch.uint8, device=self.device)
                    it = iter(layout)
                    views = tuple({k: arena[o:o + v.numel() * v.element_size()].view(dt).view(shp)
                                   for k, v in d.items() for (o, dt, shp) in (next(it),)} for d in dicts)
                    slot = self._slots[i] = (views, arena, layout)
                first = next(iter(dicts[0].values()))
                host_span = torch.empty(0, dtype=torch.uint8).set_(first.untyped_storage(), lo, (span,))
                plan = ("span", slot[1], host_span, span, dicts)
            else:
                if slot is None or slot[2] is not None or any(dv.shape != hv.shape for dd, hd in zip(slot[0], dicts)
                                                              for dv, hv in zip(dd.values(), hd.values())):
                    slot = self._slots[i] = (self._device_like(dicts), None, None)
                pairs = [(dd[k], v) for dd, hd in zip(slot[0], dicts) for k, v in hd.items()]
                plan = ("each", pairs, None, sum(v.numel() * v.element_size() for _, v in pairs), dicts)
            if len(self._plans) > 64:
                self._plans.clear()
            self._plans[(i, key)] = plan                     # holds `dicts`: the ids in the key stay valid
        with torch.cuda.stream(self.copy_stream):
            if self._free[i] is not None:
                self.copy_stream.wait_event(self._free[i])       # the consumer of the slot's previous content is done
            if plan[0] == "span":
                plan[1].copy_(plan[2], non_blocking=True)
            else:
                for dv, hv in plan[1]:
                    dv.copy_(hv, non_blocking=True)
            self._ready[i].record(self.copy_stream)
        self.bytes_per_put = plan[3]
        self._head += 1

    def get(self) -> Tuple[Dict[str, torch.Tensor], ...]:
        """Device tensors of the oldest staged step; the current stream waits for their copies."""
        if self._tail >= self._head:
            raise RuntimeError("HostFeeder: nothing staged; call put() first")
        i = self._tail % self.depth
        torch.cuda.current_stream(self.device).wait_event(self._ready[i])
        self._tail += 1
        self._handed[i] = True
        return self._slots[i][0]

    def release(self) -> None:
        """Call after the last kernel that reads the most recently handed-out slot has been enqueued."""
        i = (self._tail - 1) % self.depth
        ev = self._free[i] or torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._free[i] = ev
        self._handed[i] = False
