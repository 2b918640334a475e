"""A whole head-path step behind ONE CUDA-graph launch, fed from pinned host memory.

The reference's training loop (train.py:148-171) runs, per step, ``batch[k].to(device)``, the criterion, ``backward()``
and -- when it evaluates -- ``decode_detection`` + ``.cpu()``.  Through the plugin modules each of those is a handful of
Python calls, autograd bookkeeping and ctypes launches: ~0.3-0.4 ms of host time per step around 27 us of kernels
(tools/e2e_profile.py).  ``HostStep`` keeps the SAME calls -- ``fn`` below is ordinary code over the plugin modules
(``DetectionLoss``, ``loss.backward()``, ``decode_detection`` ...) -- but runs them once per input slot under stream capture
and replays the captured launches afterwards:

    def fn(out, batch):                                   # device tensors of one HostFeeder slot (static addresses)
        out = {k: v.detach().requires_grad_(True) for k, v in out.items()}
        work = dict(out)
        loss, stats = criterion(work, batch)
        loss.backward()
        dets = decode_detection(work['hm'], work['wh'].detach(), work['reg'].detach(), K=100)
        return {'loss': loss.detach().reshape(1), 'dets': dets, 'grad_hm': out['hm'].grad}

    step = HostStep(fn, device, fetch=('loss', 'dets'))   # results copied to pinned host memory inside the graph
    step.stage(host_out, host_batch)                      # pinned host tensors -> slot, asynchronous (copy stream)
    for ...:
        res = step.run()                                  # ONE graph launch: kernels + D2H of the fetched results
        step.stage(next_out, next_batch)                  # H2D of step i+1 rides the copy engine under step i
        res.wait(); res.host['loss'], res.host['dets']    # pinned host tensors;  res.device[...]: everything fn returned

Nothing numerical changes: the graph holds exactly the launches the eager calls make (tests/test_gpu_parity.py::
test_graphed_host_step_equals_eager_calls compares them bit for bit).  A graph is captured per distinct set of input
addresses (HostFeeder hands out ``depth`` slots, so ``depth`` graphs), after ``warmup`` eager runs that size the library's
workspaces.  What ``fn`` may not do under capture is what CUDA forbids there: synchronise, or read device results on the
host (``.item()``, ``float(t)``); ``fn`` must be a pure function of the tensors it is given (anything else it reads is
frozen at capture time).

Sharded runs: the in-kernel peer exchange (``make_sharded_loss(exchange='peers')``, what ``'auto'`` picks when the
peers' memory can be mapped) captures and replays like the single-GPU loss (tests/test_gpu_multi.py::
test_graphed_host_step_with_sharded_loss_two_gpus: bit-identical to the single-device eager calls).  The NCCL schedule
(``exchange='nccl'``: two ``all_reduce`` calls inside the step) is NOT supported here -- captured through this class it
hung both ranks of that test; run such a step eagerly.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence, Tuple

import torch

from . import _lib as L
from .feeder import HostFeeder


class StepResult:
    """Outputs of one replay.  ``device``: what ``fn`` returned (valid until the same slot runs again, i.e. for
    ``depth`` steps); ``host``: the fetched entries in pinned host memory (valid after ``wait()``, until the same
    result slot is reused ``depth`` steps later)."""

    __slots__ = ("device", "host", "_event")

    def __init__(self, device, host, event):
        self.device, self.host, self._event = device, host, event

    def done(self) -> bool:
        return self._event.query()

    def wait(self) -> "StepResult":
        self._event.synchronize()
        return self


class _Captured:
    __slots__ = ("graph", "outputs", "host", "keep")


class GraphedFn:
    """``fn(*dicts) -> dict of tensors``, captured once per distinct set of input addresses and replayed afterwards.
    ``fetch``: keys of the result copied into pinned host memory by a copy node at the end of the graph."""

    def __init__(self, fn: Callable[..., Dict[str, torch.Tensor]], device, fetch: Sequence[str] = (), warmup: int = 2,
                 max_graphs: int = 16):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("cnhead.graphed: a CUDA device is required (there is no CPU path)")
        if warmup < 1:
            raise ValueError("cnhead.graphed: at least one eager warm-up run is needed (it sizes the workspaces)")
        self.fn, self.fetch, self.warmup, self.max_graphs = fn, tuple(fetch), int(warmup), int(max_graphs)
        self._graphs: Dict[Tuple[int, ...], _Captured] = {}
        self._stream = torch.cuda.Stream(device=self.device)      # warm-up and capture run here (their own workspaces)

    @staticmethod
    def _key(dicts) -> Tuple[int, ...]:
        return tuple(v.data_ptr() for d in dicts for v in d.values())

    def _capture(self, dicts) -> _Captured:
        if len(self._graphs) >= self.max_graphs:
            raise RuntimeError(f"cnhead.graphed: more than {self.max_graphs} distinct input sets; feed the step from a "
                               "fixed set of slots (HostFeeder) instead of fresh tensors")
        cur = torch.cuda.current_stream(self.device)
        self._stream.wait_stream(cur)
        c = _Captured()
        c.host = {}
        with torch.cuda.stream(self._stream):
            for _ in range(self.warmup):
                out = self.fn(*dicts)
            if not isinstance(out, dict):
                raise RuntimeError("cnhead.graphed: fn must return a dict of tensors")
            for k in self.fetch:
                if k not in out:
                    raise RuntimeError(f"cnhead.graphed: fetch key '{k}' is not in fn's result {sorted(out)}")
                # (page-locked memory cannot be allocated under capture: sized from the warm-up run)
                c.host[k] = torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory()
            del out
        self._stream.synchronize()
        c.graph = torch.cuda.CUDAGraph()
        temps = []
        with torch.cuda.graph(c.graph, stream=self._stream):
            c.outputs = self.fn(*dicts)
            for k in self.fetch:
                t = c.outputs[k].detach()
                if not t.is_contiguous():
                    t = t.contiguous()
                if t.shape != c.host[k].shape:
                    raise RuntimeError(f"cnhead.graphed: result '{k}' changed shape between runs")
                # a copy node of the graph (the library's plain cudaMemcpyAsync; torch's non_blocking copy into pinned
                # memory also records host-allocator events on the copying stream: not something to rely on under capture)
                L.check(L.lib().cnh_copy_async(c.host[k].data_ptr(), t.data_ptr(), t.numel() * t.element_size(),
                                               L.stream_ptr()), "copy_async")
                temps.append(t)
        c.keep = (dicts, temps)                                   # the addresses in the key (and any contiguous temporaries) stay valid
        cur.wait_stream(self._stream)
        return c

    def __call__(self, *dicts: Dict[str, torch.Tensor]) -> StepResult:
        key = self._key(dicts)
        c = self._graphs.get(key)
        if c is None:
            c = self._graphs[key] = self._capture(dicts)
        c.graph.replay()                                          # on the caller's current stream
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return StepResult(c.outputs, c.host, ev)

    @property
    def n_graphs(self) -> int:
        return len(self._graphs)


class HostStep:
    """HostFeeder + GraphedFn: ``stage()`` ships one step's pinned host tensors to the next free device slot on the copy
    stream, ``run()`` replays the step's graph on the oldest staged slot.  ``resident``: dicts of device tensors handed
    to ``fn`` in front of the staged ones (head maps that never leave the GPU, as behind the reference's backbone)."""

    def __init__(self, fn, device, fetch: Sequence[str] = (), depth: int = 2, warmup: int = 2):
        self.feeder = HostFeeder(device, depth=depth)
        self.graphed = GraphedFn(fn, device, fetch=fetch, warmup=warmup, max_graphs=64)

    def stage(self, *host_dicts: Dict[str, torch.Tensor]) -> None:
        self.feeder.put(*host_dicts)

    def run(self, resident: Optional[Sequence[Dict[str, torch.Tensor]]] = None) -> StepResult:
        staged = self.feeder.get()
        res = self.graphed(*(tuple(resident or ()) + tuple(staged)))
        self.feeder.release()
        return res
