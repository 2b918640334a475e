"""The e2e legs of bench.py in a chosen order (is a leg's number its own, or its position's?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "centernet-uda_b200"), ROOT]
import torch
import bench
from cnhead import synthetic
cfg = synthetic.CONFIGS["cfg2"]
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
order = sys.argv[1:] or ["boxes", "dense", "targets_only", "dense:eager", "dense", "boxes"]
for m in order:
    mode, _, how = m.partition(":")
    e = bench.run_e2e(cfg, 16, 0, 1, dev, 400, 10, mode, how != "eager")
    print(f"{m:22s} {e['value']:10.0f} heatmaps/s  {e['ms_per_step']*1e3:7.1f} us/step  {e['h2d_GBps']:5.1f} GB/s  {e['h2d_bytes_per_step']} B")
