"""decode launch alone (graph replay over rotating buffer sets, CUDA events): us per launch and fraction of the HBM
peak per config, plus the cluster size the library picked.  CNH_DECODE_CS caps the cluster size.
Usage: python tools/dec_time.py [cfg2 cfg5 ...]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "centernet-uda_b200"), ROOT]
import torch
import bench
from cnhead import _lib as L, synthetic

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
peak, _ = bench.hbm_peak()
lib = L.lib()
lib.cnh_debug_active_clusters.argtypes = [C.c_int]
if os.environ.get("CNH_DECODE_CS") is None:
    print("co-resident clusters per size:", {cs: lib.cnh_debug_active_clusters(cs) for cs in range(1, 9)})
out = {}
for name in sys.argv[1:] or ["cfg2", "cfg5"]:
    cfg = synthetic.CONFIGS[name]
    batch = cfg.batch if name != "cfg5" else 16
    w = bench.Workload(cfg, batch, 0, 1, dev, True)
    lib.cnh_debug_decode_cluster.argtypes = [C.c_void_p]
    cs = lib.cnh_debug_decode_cluster(C.byref(w.dstep.dec_args[0])) // 1000
    us = w.time("decode_only", 200, 20, False) * 1e3
    step = w.time("step", 200, 20, False) * 1e3
    out[name] = {"cluster": cs, "decode_us": round(us, 2), "frac": round(w.dec_bytes / (us * 1e-6) / 1e9 / peak, 3),
                 "step_us": round(step, 2), "step_frac": round(w.step_bytes / (step * 1e-6) / 1e9 / peak, 3)}
    w.close()
print("CNH_DECODE_CS=%s" % os.environ.get("CNH_DECODE_CS"), json.dumps(out))
