"""Where a step's time goes: CUDA events after every launch of a cfg5 (or given) step, eager launches,
averaged.  Usage: python tools/step_events.py [cfg5] [--fuse]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "centernet-uda_b200"), ROOT]
import torch
import bench
from cnhead import _lib as L, synthetic

name = next((a for a in sys.argv[1:] if not a.startswith("-")), "cfg5")
bench.DeviceStep.FUSE = "--fuse" in sys.argv
cfg = synthetic.CONFIGS[name]
batch = cfg.batch if name != "cfg5" else 16
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
w = bench.Workload(cfg, batch, 0, 1, dev, False)
d = w.dstep
lib = d.lib
st = torch.cuda.Stream()
N = 60
acc = {}
with torch.cuda.stream(st):
    for it in range(N + 10):
        i = it % w.n_sets
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        sp = L.stream_ptr()
        ev[0].record()
        L.check(lib.cnh_detloss_fused(C.byref(d.loss_args[i]), d.ws_loss.data_ptr(), d.ws_loss.numel(), sp), "loss")
        ev[1].record()
        L.check(lib.cnh_scale_inplace(C.byref(d.scale_args[i]), sp), "scale")
        ev[2].record()
        d.decode_step(i)
        ev[3].record()
        torch.cuda.synchronize()
        if it >= 10:
            for k, (a, b) in {"loss": (0, 1), "scale": (1, 2), "decode": (2, 3), "step": (0, 3)}.items():
                acc[k] = acc.get(k, 0.0) + ev[a].elapsed_time(ev[b]) * 1e3
print(name, "fused" if d.fused_decode() else "unfused", {k: round(v / N, 2) for k, v in acc.items()}, "us (eager, one step at a time)")
