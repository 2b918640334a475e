import ast
import glob
import os
import sys

import numpy as np
import pytest
import torch

os.environ.setdefault("CNH_DECODE_ENV_RELOAD", "1")   # the decode tests flip the library's launch-shape switches per case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "centernet-uda_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_head_case(g):
    """Split a detloss_* fixture into (output dict, batch dict, ctor kwargs, grad_scale)."""
    out = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("in_")}
    bt = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("bt_")}
    kw = {str(n): ast.literal_eval(str(v)) for n, v in zip(g["kw_names"], g["kw_vals"])}
    return out, bt, kw, float(g["grad_scale"])


def rel_err(a, b):
    """normwise relative error max|a-b| / max|b| (the parity metric for fp32 tensors)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    denom = b.abs().max().item()
    if denom == 0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / denom
