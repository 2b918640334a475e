"""Restatement of the reference's CALLERS of the head path -- the training steps of ``uda/*.py`` -- with the
loss / decode modules injected, so that the very same statements can be driven by the reference's own modules
(CPU, to pin this restatement: tests/test_callers.py, where /root/reference exists), by the oracle, or by the
B200 plugin modules (GPU parity of whole steps).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function cites the reference lines it follows
(paths relative to the reference checkout); statements are kept in the reference's order, including its quirks
(in-place ``*=`` / ``/=`` on returned losses, ``adversarial_entropy_minimization.py:116``: the source map handed
to the discriminator is the ALREADY SIGMOIDED tensor that DetectionLoss rebound into ``outputs['hm']``).
"""
from __future__ import annotations

import types
from typing import Callable, Dict

import numpy as np
import torch
from torch import nn


# --------------------------------------------------------------------------------------------- #
# a deterministic miniature of the model around the head path
# --------------------------------------------------------------------------------------------- #
class TinyBackend(nn.Module):
    """Two-layer stand-in for ``backends/dla.py``: stride-4 features, then one 1x1 conv per head
    (``hm`` bias -2.19 as backends/dla.py:485).  ``down_ratio`` as backends/dla.py:513-514."""

    down_ratio = 4

    def __init__(self, num_classes: int, width: int = 8):
        super().__init__()
        self.stem = nn.Conv2d(3, width, 3, stride=4, padding=1)
        self.hm = nn.Conv2d(width, num_classes, 1)
        self.wh = nn.Conv2d(width, 2, 1)
        self.reg = nn.Conv2d(width, 2, 1)
        nn.init.constant_(self.hm.bias, -2.19)

    def forward(self, x):
        f = torch.relu(self.stem(x))
        return {"hm": self.hm(f), "wh": self.wh(f), "reg": self.reg(f)}


def fc_discriminator(num_classes: int, ndf: int = 64) -> nn.Module:
    """uda/adversarial_entropy_minimization.py:55-72 ``get_fc_discriminator`` (five stride-2 convolutions)."""
    return nn.Sequential(
        nn.Conv2d(num_classes, ndf, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(negative_slope=0.2, inplace=True),
        nn.Conv2d(ndf, ndf * 2, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(negative_slope=0.2, inplace=True),
        nn.Conv2d(ndf * 2, ndf * 4, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(negative_slope=0.2, inplace=True),
        nn.Conv2d(ndf * 4, ndf * 8, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(negative_slope=0.2, inplace=True),
        nn.Conv2d(ndf * 8, 1, kernel_size=4, stride=2, padding=1))


def tiny_case(seed: int = 7, batch: int = 2, num_classes: int = 3, size: int = 128, max_objects: int = 20):
    """Inputs, targets and freshly initialised modules of one miniature training step, all from ``seed`` (CPU
    generator: identical wherever this torch build runs).  Returns (data dict of CPU tensors, backend, discriminator)."""
    from . import head
    g = torch.Generator().manual_seed(seed)
    hw = size // TinyBackend.down_ratio
    rng = np.random.RandomState(seed)
    boxes = np.zeros((batch, max_objects, 4), dtype=np.float32)
    classes = np.zeros((batch, max_objects), dtype=np.int32)
    n_obj = rng.randint(2, 7, size=batch).astype(np.int32)
    for b in range(batch):
        for k in range(int(n_obj[b])):
            cx, cy = rng.uniform(3, hw - 3, size=2)
            w, h = rng.uniform(2, hw / 2, size=2)
            boxes[b, k] = (cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2)
            classes[b, k] = rng.randint(0, num_classes)
    tgt = head.raster_targets(boxes, classes, n_obj, num_classes, hw, hw)
    data = {k: torch.from_numpy(v) for k, v in tgt.items()}
    data["input"] = torch.randn(batch, 3, size, size, generator=g)
    data["target_domain_input"] = torch.randn(batch, 3, size, size, generator=g) * 1.3 + 0.2
    data["id"] = torch.arange(batch)
    gt = torch.zeros(batch, max_objects, 6)
    gt[..., :4] = torch.from_numpy(boxes)
    gt[..., 4] = 1.0
    gt[..., 5] = torch.from_numpy(classes).float()
    data["gt_dets"] = gt
    data["gt_areas"] = (gt[..., 2] - gt[..., 0]) * (gt[..., 3] - gt[..., 1])
    torch.manual_seed(seed)
    backend = TinyBackend(num_classes)
    disc = fc_discriminator(num_classes, ndf=8)
    return data, backend, disc


def clone_data(data: Dict[str, torch.Tensor], device=None) -> Dict[str, torch.Tensor]:
    return {k: (v.clone() if device is None else v.to(device)) for k, v in data.items()}


# --------------------------------------------------------------------------------------------- #
# uda/base.py
# --------------------------------------------------------------------------------------------- #
def base_step(backend, optimizer, centernet_loss, data, is_training=True):
    """uda/base.py:31-56 ``Model.step`` with ``criterion`` of :70-71."""
    if is_training:
        optimizer.zero_grad()                                          # :36
    outputs_source_domain = backend(data["input"])                     # :38
    outputs = {"source_domain": outputs_source_domain}                 # :40-42
    loss, stats = centernet_loss(outputs["source_domain"], data)       # :43, :70-71
    if is_training:
        loss.backward()                                                # :46
        optimizer.step()                                               # :47
    stats["total_loss"] = loss                                         # :49
    for s in stats:
        stats[s] = stats[s].cpu().detach()                             # :51-52
    outputs["stats"] = stats
    return outputs


def get_detections(outputs, batch, decode_detection: Callable, max_detections: int, rotated: bool, down_ratio: int):
    """uda/base.py:73-139 ``Model.get_detections`` (prediction half and the ground-truth bookkeeping)."""
    src = outputs["source_domain"]
    dets = decode_detection(src["hm"], src["wh"], src["reg"], kps=src["kps"] if "kps" in src else None,
                            K=max_detections, rotated=rotated)        # :76-82
    dets = dets.detach().cpu().numpy()                                 # :89
    dets[:, :, :4] *= down_ratio                                       # :90
    ids = batch["id"].cpu().numpy()                                    # :92
    mask = (batch["reg_mask"].detach().cpu().numpy() == 1).squeeze()   # :93
    dets_gt = batch["gt_dets"].cpu().numpy().copy()                    # :94
    areas_gt = batch["gt_areas"].cpu().numpy()
    dets_gt[:, :, :4] *= down_ratio                                    # :96
    box_idx, cls_idx = (5, 6) if rotated else (4, 5)                   # :107-112
    gt_boxes, gt_clss, gt_ids, gt_areas = [], [], [], []
    for i in range(dets_gt.shape[0]):                                  # :114-122
        det_gt = dets_gt[i, mask[i]]
        gt_boxes.append(det_gt[:, :box_idx])
        gt_clss.append(det_gt[:, cls_idx].astype(np.int32))
        gt_ids.append(ids[i])
        gt_areas.append(areas_gt[i, mask[i]])
    return {"pred_boxes": dets[:, :, :box_idx], "pred_classes": dets[:, :, cls_idx].astype(np.int32),
            "pred_scores": dets[:, :, box_idx], "gt_boxes": gt_boxes, "gt_classes": gt_clss, "gt_ids": gt_ids,
            "gt_areas": gt_areas}                                      # :124-132


# --------------------------------------------------------------------------------------------- #
# uda/entropy_minimization.py, uda/max_squares_minimization.py
# --------------------------------------------------------------------------------------------- #
def entropy_minimization_step(backend, optimizer, centernet_loss, entropy_loss, entropy_weight, data, is_training=True):
    """uda/entropy_minimization.py:11-43 ``EntropyMinimization.step``."""
    if is_training:
        optimizer.zero_grad()
    outputs_source_domain = backend(data["input"])                     # :18
    outputs_target_domain = backend(data["target_domain_input"])       # :19
    outputs = {"source_domain": outputs_source_domain, "target_domain": outputs_target_domain}
    c_loss, c_stats = centernet_loss(outputs["source_domain"], data)   # :26
    e_loss, e_stats = entropy_loss(outputs["target_domain"], data)     # :27
    e_loss *= entropy_weight                                           # :28 (in place, on the returned tensor)
    if is_training:
        c_loss.backward()                                              # :31
        e_loss.backward()                                              # :32
        optimizer.step()
    stats = {**c_stats, **e_stats}
    stats["total_loss"] = c_loss + e_loss                              # :36
    for s in stats:
        stats[s] = stats[s].cpu().detach()
    outputs["stats"] = stats
    return outputs


def max_squares_step(backend, optimizer, centernet_loss, max_squares_loss, max_squares_weight, data, is_training=True):
    """uda/max_squares_minimization.py:12-52 ``MaxSquaresMinimization.criterion`` + ``.step``."""
    if is_training:
        optimizer.zero_grad()
    outputs = {"source_domain": backend(data["input"]), "target_domain": backend(data["target_domain_input"])}
    s_loss, s_stats = centernet_loss(outputs["source_domain"], data)   # :13
    t_loss, t_stats = max_squares_loss(outputs["target_domain"], data)  # :14-15
    t_loss *= max_squares_weight                                       # :16
    stats = {**s_stats, **t_stats}
    if is_training:
        s_loss.backward()                                              # :40
        t_loss.backward()                                              # :41
        optimizer.step()
    stats["total_loss"] = s_loss + t_loss                              # :44
    for s in stats:
        stats[s] = stats[s].cpu().detach()
    outputs["stats"] = stats
    return outputs


# --------------------------------------------------------------------------------------------- #
# uda/adversarial_entropy_minimization.py
# --------------------------------------------------------------------------------------------- #
def advent_step(backend, optimizer, discriminator, discriminator_optimizer, centernet_loss, adversarial_loss,
                entropy_map: Callable, adversarial_weight: float, data, is_training=True,
                source_label: int = 0, target_label: int = 1):
    """uda/adversarial_entropy_minimization.py:77-152 ``AdversarialEntropyMinimization.step`` (ADVENT)."""
    if is_training:
        optimizer.zero_grad()                                          # :82
        discriminator_optimizer.zero_grad()                            # :83
    for param in discriminator.parameters():
        param.requires_grad = False                                    # :85-86
    outputs_source_domain = backend(data["input"])                     # :88
    outputs_target_domain = backend(data["target_domain_input"])       # :89
    outputs_target_generator = discriminator(entropy_map(outputs_target_domain["hm"]))   # :91-92
    outputs = {"source_domain": outputs_source_domain, "target_domain": outputs_target_domain}
    loss, stats = centernet_loss(outputs_source_domain, data)          # :99
    if is_training:
        loss.backward()                                                # :101
    dtf_loss, dtf_stats = adversarial_loss(outputs_target_generator, source_label)   # :104-106  fool the discriminator
    dtf_loss *= adversarial_weight                                     # :107
    if is_training:
        dtf_loss.backward()                                            # :110  (through the map into the backbone)
    for param in discriminator.parameters():
        param.requires_grad = True                                     # :113-114
    source = outputs_source_domain["hm"].detach()                      # :116  ALREADY sigmoided + clamped (rebound by the loss)
    target = outputs_target_domain["hm"].detach()                      # :117  raw logits
    outputs_source_generator = discriminator(entropy_map(source))      # :119
    ds_loss, ds_stats = adversarial_loss(outputs_source_generator, source_label)      # :120-121
    ds_loss /= 2.0                                                     # :122
    if is_training:
        ds_loss.backward()                                             # :125
    outputs_target_generator = discriminator(entropy_map(target))      # :127
    dt_loss, dt_stats = adversarial_loss(outputs_target_generator, target_label)      # :128-129
    dt_loss /= 2.0                                                     # :130
    if is_training:
        dt_loss.backward()                                             # :133
    outputs["source_generator"] = outputs_source_generator             # :135
    outputs["target_generator"] = outputs_target_domain                # :136 (sic)
    if is_training:
        optimizer.step()                                               # :139
        discriminator_optimizer.step()                                 # :140
    stats["total_loss"] = loss + ds_loss + dt_loss + dtf_loss          # :143
    stats["dis_soruce"] = ds_loss                                      # :144 (sic)
    stats["dis_target"] = dt_loss
    stats["dis_fool"] = dtf_loss
    for s in stats:
        stats[s] = stats[s].cpu().detach()
    outputs["stats"] = stats
    return outputs


class CpuAdventLoss(nn.Module):
    """losses/advent.py:5-18 for tensors on ANY device: the reference builds its label with
    ``.to(y_pred.get_device())``, which raises for CPU tensors (get_device() == -1), so the CPU runs of
    the ADVENT step state what :8 and :16 compute."""

    def __init__(self):
        super().__init__()
        self.crit = nn.BCEWithLogitsLoss()

    def forward(self, y_pred, y_true):
        advent_loss = self.crit(y_pred, torch.full_like(y_pred, float(y_true)))
        return advent_loss, {"advent_loss": advent_loss}


def grads_of(module: nn.Module) -> Dict[str, np.ndarray]:
    return {n: (p.grad.detach().cpu().numpy().copy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
            for n, p in module.named_parameters()}


def reference_stubs():
    """``hydra`` and ``omegaconf`` are not installed here; the reference's ``utils/helper.py`` and
    ``uda/adversarial_entropy_minimization.py`` import them at module level.  Returns sys.modules entries that are
    enough to import (and run the steps of) the unmodified ``uda`` package."""
    import importlib
    hydra = types.ModuleType("hydra")
    hydra.utils = types.ModuleType("hydra.utils")

    def get_class(path):
        mod, _, name = path.rpartition(".")
        return getattr(importlib.import_module(mod), name)

    hydra.utils.get_class = get_class
    hydra.utils.get_method = get_class
    omegaconf = types.ModuleType("omegaconf")
    listconfig = types.ModuleType("omegaconf.listconfig")
    listconfig.ListConfig = list
    omegaconf.listconfig = listconfig
    return {"hydra": hydra, "hydra.utils": hydra.utils, "omegaconf": omegaconf, "omegaconf.listconfig": listconfig}


# --------------------------------------------------------------------------------------------- #
# the oracle's functions behind the reference's module interfaces (what the callers above are handed on CPU)
# --------------------------------------------------------------------------------------------- #
class OracleDetectionLoss(nn.Module):
    """``crit(output, batch) -> (loss, stats)`` with the ``output['hm']`` rebinding of losses/centernet.py:34."""

    def __init__(self, **kw):
        super().__init__()
        self.kw = kw

    def forward(self, output, batch):
        from . import head
        loss, stats, prob = head.detection_loss(output, batch, **self.kw)
        output["hm"] = prob
        stats = {k: stats[k] for k in ("centernet_loss", "hm_loss", "wh_loss", "off_loss", "kp_loss") if k in stats}
        return loss, stats


class OracleEntropyLoss(nn.Module):
    def forward(self, outputs, batch):
        from . import head
        loss = head.entropy_loss(outputs["hm"])
        return loss, {"entropy_loss": loss}


class OracleMaxSquareLoss(nn.Module):
    def forward(self, outputs, batch):
        from . import head
        loss = head.max_square_loss(outputs["hm"])
        return loss, {"max_square_loss": loss}


def oracle_decode_detection(heat, wh, reg=None, kps=None, K=100, rotated=False):
    from . import head
    return head.decode_stable(heat, wh, reg, kps, K=K, rotated=rotated)[0]
