#!/usr/bin/env python
"""Generate tests/golden/callers_*.npz by running the UNMODIFIED callers of the head path -- ``uda/base.py``,
``uda/entropy_minimization.py``, ``uda/max_squares_minimization.py``, ``uda/adversarial_entropy_minimization.py``
of the reference -- on CPU with the reference's own loss / decode modules, around the miniature model of
``oracle/callers.py::tiny_case`` (inputs, targets and initial weights come from a seed, so the fixtures hold
results only).  Run in the build container (needs /root/reference):

    python tests/golden/make_callers_golden.py

``hydra`` / ``omegaconf`` are stubbed (``oracle.callers.reference_stubs``: not installed, imported at module level
by the reference).  The one substitution: ``losses.advent.AdventLoss`` cannot run on CPU tensors
(losses/advent.py:14), so the ADVENT step uses ``oracle.callers.CpuAdventLoss``, which states :8,:16.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = os.environ.get("CNH_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import callers                              # noqa: E402

sys.modules.update(callers.reference_stubs())
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")
torch.set_num_threads(1)

from losses.centernet import DetectionLoss              # noqa: E402  (the reference's)
from uda.base import Model                              # noqa: E402
from uda.entropy_minimization import EntropyMinimization                      # noqa: E402
from uda.max_squares_minimization import MaxSquaresMinimization               # noqa: E402
from uda.adversarial_entropy_minimization import AdversarialEntropyMinimization  # noqa: E402

LOSS_KW = dict(hm_weight=1.0, wh_weight=0.1, off_weight=1.0)
K = 20


def cfg():
    params = types.SimpleNamespace(rotated_boxes=False, num_classes=3)
    return types.SimpleNamespace(max_detections=K, model=types.SimpleNamespace(backend=types.SimpleNamespace(params=params)))


def wire(model, backend):
    model.cfg = cfg()
    model.backend = backend
    model.device = torch.device("cpu")
    model.optimizer = torch.optim.SGD(backend.parameters(), lr=0.01)
    model.centernet_loss = DetectionLoss(**LOSS_KW)
    return model


def pack(prefix, outputs, backend, extra=None):
    out = {f"{prefix}stat_{k}": v.numpy() for k, v in outputs["stats"].items()}
    out.update({f"{prefix}grad_{k}": v for k, v in callers.grads_of(backend).items()})
    out[f"{prefix}prob"] = outputs["source_domain"]["hm"].detach().numpy()
    if extra:
        out.update({prefix + k: v for k, v in extra.items()})
    return out


def main():
    res = {}
    data, backend, disc = callers.tiny_case()
    res["input_checksum"] = np.array([float(data["input"].double().sum()), float(data["target_domain_input"].double().sum()),
                                      float(sum(p.double().sum() for p in backend.parameters()))])
    # uda/base.py: step + get_detections
    m = wire(Model(), backend)
    d = callers.clone_data(data)
    outputs = m.step(d, is_training=True)
    dets = m.get_detections(outputs, d)
    res.update(pack("base_", outputs, backend, {"pred_boxes": dets["pred_boxes"], "pred_scores": dets["pred_scores"],
                                                "pred_classes": dets["pred_classes"],
                                                "gt_boxes0": dets["gt_boxes"][0], "gt_classes0": dets["gt_classes"][0]}))
    # uda/entropy_minimization.py
    data, backend, disc = callers.tiny_case()
    m = wire(EntropyMinimization(entropy_weight=0.1), backend)
    res.update(pack("ent_", m.step(callers.clone_data(data), is_training=True), backend))
    # uda/max_squares_minimization.py
    data, backend, disc = callers.tiny_case()
    m = wire(MaxSquaresMinimization(max_squares_weight=0.2), backend)
    res.update(pack("msq_", m.step(callers.clone_data(data), is_training=True), backend))
    # uda/adversarial_entropy_minimization.py
    data, backend, disc = callers.tiny_case()
    m = wire(AdversarialEntropyMinimization(adversarial_weight=0.01), backend)
    m.adversarial_loss = callers.CpuAdventLoss()
    m.discriminator = disc
    m.discriminator_optimizer = torch.optim.Adam(disc.parameters())
    m.discriminator_scheduler = None
    outputs = m.step(callers.clone_data(data), is_training=True)
    res.update(pack("adv_", outputs, backend, {"dgrad_" + k: v for k, v in callers.grads_of(disc).items()}))
    np.savez_compressed(os.path.join(HERE, "callers_steps.npz"), **res)
    print("wrote callers_steps.npz:", len(res), "arrays;", {k: float(v) for k, v in res.items() if k.startswith("adv_stat_")})


if __name__ == "__main__":
    main()
