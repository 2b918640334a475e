"""CPU restatement of the reference's per-pixel head algorithms (torch on CPU).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The reference's arithmetic for this path lives in PyTorch ATen ops (SURVEY.md
section 8c), so the faithful restatement is the same sequence of ATen ops on CPU
tensors, written functionally (no modules, no in-place side effects on the
caller's tensors).  Every function cites the reference file:line it follows
(paths relative to the reference checkout).  All functions accept float32 (the
reference's precision; parity target) or float64 (used by the tests as a
higher-precision truth when judging *which* of two fp32 results is closer).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

LO = 1e-4          # utils/tensor.py:6  clamp bounds
HI = 1 - 1e-4


# --------------------------------------------------------------------------- #
# utils/tensor.py
# --------------------------------------------------------------------------- #
def sigmoid_clamp(x: torch.Tensor) -> torch.Tensor:
    """utils/tensor.py:5-7 ``_sigmoid``: clamp(sigmoid(x), 1e-4, 1-1e-4).

    Out of place (the reference overwrites ``x`` with the unclamped sigmoid; no
    caller reads that side effect)."""
    return torch.clamp(torch.sigmoid(x), min=LO, max=HI)


def gather_rows(fmap: torch.Tensor, ind: torch.Tensor) -> torch.Tensor:
    """utils/tensor.py:10-25 ``_transpose_and_gather_feat``.

    fmap [B,D,H,W], ind [B,M] int64 (flat y*W+x)  ->  [B,M,D]."""
    b, d = fmap.shape[:2]
    rows = fmap.reshape(b, d, -1).transpose(1, 2)            # [B,HW,D]
    return torch.gather(rows, 1, ind.unsqueeze(-1).expand(-1, -1, d))


# --------------------------------------------------------------------------- #
# losses/centernet.py
# --------------------------------------------------------------------------- #
def focal_terms(prob: torch.Tensor, gt: torch.Tensor):
    """losses/centernet.py:76-89: (pos_sum, neg_sum, num_pos) over the whole batch."""
    is_pos = (gt == 1).to(prob.dtype)
    is_neg = (gt < 1).to(prob.dtype)
    pos = torch.log(prob) * (1 - prob) ** 2 * is_pos
    neg = torch.log(1 - prob) * prob ** 2 * (1 - gt) ** 4 * is_neg
    return pos.sum(), neg.sum(), is_pos.sum()


def focal_loss(prob: torch.Tensor, gt: torch.Tensor, weight: float = 1.0) -> torch.Tensor:
    """losses/centernet.py:69-95 ``FocalLoss._neg_loss`` (penalty-reduced focal)."""
    pos_sum, neg_sum, num_pos = focal_terms(prob, gt)
    if float(num_pos) == 0:                                   # :91-92
        return -neg_sum * weight
    return -(pos_sum + neg_sum) / num_pos * weight            # :94-95


def _expanded_mask(mask: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return mask.unsqueeze(2).expand_as(like).to(like.dtype)


def masked_l1(fmap, mask, ind, target, weight=1.0, angle_weight=1.0) -> torch.Tensor:
    """losses/centernet.py:98-133 ``RegL1Loss.forward``.

    D != 3: weight * sum|pred*m - tgt*m| / (sum(m_expanded) + 1e-4).
    D == 3: wh part on [...,0:2] * weight  +  angle_weight * sum|sc(pred_a*m) -
    sc(tgt_a*m)| / same denominator (sc = sigmoid_clamp, applied to the *target*
    too, :114-117)."""
    pred = gather_rows(fmap, ind)
    m = _expanded_mask(mask, pred)
    pred = pred * m
    tgt = target.to(pred.dtype) * m
    denom = m.sum() + 1e-4
    if pred.shape[-1] == 3:
        wh = (pred[..., 0:2] - tgt[..., 0:2]).abs().sum() / denom
        ang = (sigmoid_clamp(pred[..., 2:3]) - sigmoid_clamp(tgt[..., 2:3])).abs().sum() / denom
        return wh * weight + ang * angle_weight
    return (pred - tgt).abs().sum() / denom * weight


def periodic_l1(fmap, mask, ind, target, wh_weight=1.0, angle_weight=1.0) -> torch.Tensor:
    """losses/centernet.py:192-223 ``PeriodicRegL1Loss.forward`` (RAPiD periodic L1).

    angle prediction sc(pred)*2pi - pi (radians), target in degrees -> deg2rad;
    |remainder((pa - ta) - pi/2, pi) - pi/2| summed / (sum(m_expanded)+1e-4)."""
    pred = gather_rows(fmap, ind)
    m = _expanded_mask(mask, pred)
    pred = pred * m
    tgt = target.to(pred.dtype) * m
    denom = m.sum() + 1e-4
    wh = (pred[..., 0:2] - tgt[..., 0:2]).abs().sum() / denom
    pa = sigmoid_clamp(pred[..., 2:3]) * 2 * math.pi - math.pi
    ta = torch.deg2rad(tgt[..., 2:3])
    per = (torch.remainder((pa - ta) - math.pi / 2, math.pi) - math.pi / 2).abs().sum() / denom
    return wh * wh_weight + per * angle_weight


def keypoint_l1(fmap, mask, ind, target, weight=1.0, pair_indices=None,
                distance_weight=0.1, use_l1=False) -> torch.Tensor:
    """losses/centernet.py:136-189 ``KPSL1Loss.forward`` (element-wise mask [B,M,2nk];
    optional limb-length term with the literal +1e4 under the sqrt, :177-178)."""
    pred = gather_rows(fmap, ind)
    m = mask.to(pred.dtype)
    pred = pred * m
    tgt = target.to(pred.dtype) * m
    denom = m.sum() + 1e-4
    loss = (pred - tgt).abs().sum() / denom * weight
    if pair_indices is not None:
        pairs = torch.as_tensor(pair_indices)
        n, c, k2 = tgt.shape
        p = pred.reshape(n, c, k2 // 2, 2)
        t = tgt.reshape(n, c, k2 // 2, 2)
        pa, pb = p[:, :, pairs[:, 0]], p[:, :, pairs[:, 1]]
        ta, tb = t[:, :, pairs[:, 0]], t[:, :, pairs[:, 1]]
        if use_l1:
            dp, dt = (pa - pb).abs().sum(-1), (ta - tb).abs().sum(-1)
        else:
            dp = (((pa - pb) ** 2).sum(-1) + 1e4) ** 0.5
            dt = (((ta - tb) ** 2).sum(-1) + 1e4) ** 0.5
        loss = loss + (dp - dt).abs().sum() / denom * distance_weight
    return loss


def detection_loss(output: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor],
                   hm_weight=1.0, wh_weight=0.1, off_weight=1.0, angle_weight=1.0,
                   periodic=False, kp_weight=None, kp_indices=None,
                   kp_distance_weight=0.1, kp_distance_weight_l1=False
                   ) -> Tuple[torch.Tensor, Dict[str, torch.Tensor], torch.Tensor]:
    """losses/centernet.py:7-56 ``DetectionLoss``: returns (loss, stats, clamped prob).

    ``prob`` is what the reference rebinds into ``output['hm']`` (:34)."""
    prob = sigmoid_clamp(output["hm"])
    hm = focal_loss(prob, batch["hm"].to(prob.dtype), hm_weight)
    if periodic:
        wh = periodic_l1(output["wh"], batch["reg_mask"], batch["ind"], batch["wh"],
                         wh_weight, angle_weight)
    else:
        wh = masked_l1(output["wh"], batch["reg_mask"], batch["ind"], batch["wh"],
                       wh_weight, angle_weight)
    off = masked_l1(output["reg"], batch["reg_mask"], batch["ind"], batch["reg"], off_weight)
    loss = hm + wh + off
    stats = {"hm_loss": hm, "wh_loss": wh, "off_loss": off}
    if kp_weight is not None or kp_indices is not None:
        kp = keypoint_l1(output["kps"], batch["kp_reg_mask"], batch["ind"], batch["kps"],
                         kp_weight, kp_indices, kp_distance_weight, kp_distance_weight_l1)
        loss = loss + kp
        stats["kp_loss"] = kp
    stats["centernet_loss"] = loss
    return loss, stats, prob


def detection_loss_with_grads(output, batch, grad_scale: float = 1.0, **cfg):
    """fwd + autograd bwd exactly as ``uda/base.py:43-46`` drives it; returns
    (loss, stats, prob, {head: dLoss/dhead}) without touching the caller's tensors."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in output.items()}
    tgt = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    loss, stats, prob = detection_loss(leaves, tgt, **cfg)
    (loss * grad_scale).backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return loss.detach(), {k: s.detach() for k, s in stats.items()}, prob.detach(), grads


# --------------------------------------------------------------------------- #
# backends/decode.py
# --------------------------------------------------------------------------- #
def peak_scores(heat: torch.Tensor, kernel: int = 3) -> torch.Tensor:
    """backends/decode.py:6-13 ``_nms``: heat * (1 - ceil(maxpool3x3(heat) - heat))."""
    pad = (kernel - 1) // 2
    hmax = torch.nn.functional.max_pool2d(heat, (kernel, kernel), stride=1, padding=pad)
    return heat * (1.0 - torch.ceil(hmax - heat))


def decode_stable(heat, wh, reg=None, kps=None, K=100, rotated=False, nms_size=3):
    """backends/decode.py:35-76 ``decode_detection`` with the tie rule made explicit.

    The reference's two-stage ``torch.topk`` (:16-32) equals a per-sample top-K over
    C*H*W whose tie order torch leaves unspecified; the oracle fixes it: stable
    descending sort of the NMS map on flat index c*HW + y*W + x (lower index wins).
    Returns (detections [B,K,6|7], flat_index [B,K] int64[, kps [B,K,nk,2]])."""
    b, c, h, w = heat.shape
    scores = peak_scores(heat, nms_size).reshape(b, -1)
    order = torch.sort(scores, dim=1, descending=True, stable=True).indices[:, :K]
    top = torch.gather(scores, 1, order)
    cls = (order // (h * w)).to(heat.dtype)
    pix = order % (h * w)
    ys = (pix // w).to(heat.dtype).unsqueeze(-1)
    xs = (pix % w).to(heat.dtype).unsqueeze(-1)
    if reg is not None:                                       # :44-48
        off = gather_rows(reg, pix)
        xs = xs + off[..., 0:1]
        ys = ys + off[..., 1:2]
    else:                                                     # :49-51
        xs = xs + 0.5
        ys = ys + 0.5
    size = gather_rows(wh, pix)
    if not rotated:                                           # :56-60
        box = torch.cat([xs - size[..., 0:1] / 2, ys - size[..., 1:2] / 2,
                         xs + size[..., 0:1] / 2, ys + size[..., 1:2] / 2], dim=2)
    else:                                                     # :61-66
        box = torch.cat([xs, ys, size[..., 0:1], size[..., 1:2],
                         sigmoid_clamp(size[..., 2:3]) * 360.0 - 180.0], dim=2)
    dets = torch.cat([box, top.unsqueeze(-1), cls.unsqueeze(-1)], dim=2)
    if kps is not None:                                       # :69-74
        pts = gather_rows(kps, pix)
        pts = pts.reshape(b, K, pts.shape[2] // 2, 2).clone()
        pts[..., 0] += xs
        pts[..., 1] += ys
        return dets, order, pts
    return dets, order


def decode_two_stage(heat, wh, reg=None, K=100, rotated=False):
    """backends/decode.py:16-76 with the reference's own selection structure: top-K per class over
    H*W, then top-K over the C*K survivors (torch.topk: tie order unspecified).  Used as the timed
    CPU baseline (bench.py); parity is judged against ``decode_stable``."""
    b, c, h, w = heat.shape
    nms = peak_scores(heat)
    s1, i1 = torch.topk(nms.reshape(b, c, -1), K)                       # :19
    s2, i2 = torch.topk(s1.reshape(b, -1), K)                            # :25
    cls = (i2 // K).to(heat.dtype)                                      # :26
    pix = torch.gather(i1.reshape(b, -1), 1, i2)                        # :27-28
    ys = (pix // w).to(heat.dtype).unsqueeze(-1)
    xs = (pix % w).to(heat.dtype).unsqueeze(-1)
    if reg is not None:
        off = gather_rows(reg, pix)
        xs, ys = xs + off[..., 0:1], ys + off[..., 1:2]
    else:
        xs, ys = xs + 0.5, ys + 0.5
    size = gather_rows(wh, pix)
    if not rotated:
        box = torch.cat([xs - size[..., 0:1] / 2, ys - size[..., 1:2] / 2,
                         xs + size[..., 0:1] / 2, ys + size[..., 1:2] / 2], dim=2)
    else:
        box = torch.cat([xs, ys, size[..., 0:1], size[..., 1:2],
                         sigmoid_clamp(size[..., 2:3]) * 360.0 - 180.0], dim=2)
    return torch.cat([box, s2.unsqueeze(-1), cls.unsqueeze(-1)], dim=2)


# --------------------------------------------------------------------------- #
# losses/entropy.py, losses/max_square.py, losses/advent.py, utils/image.py
# --------------------------------------------------------------------------- #
def entropy_loss(logits: torch.Tensor, eta: Optional[float] = None) -> torch.Tensor:
    """losses/entropy.py:10-28: Shannon entropy of the channel softmax, normalised by
    N*H*W*log2(C) (log2 C evaluated in fp32, :25); FDA robust variant when eta is set."""
    v = torch.softmax(logits, dim=1)
    n, c, h, w = v.shape
    log2c = torch.log2(torch.tensor([float(c)], dtype=torch.float32)).to(v.dtype)
    if eta is not None:                                       # :18-22
        ent = -(v * torch.log2(v + 1e-30)).sum(dim=1) / log2c
        return ((ent ** 2.0 + 1e-30) ** eta).mean()
    return (-(v * torch.log2(v + 1e-30)).sum() / (n * h * w * log2c)).squeeze()


def max_square_loss(logits: torch.Tensor) -> torch.Tensor:
    """losses/max_square.py:6-14: -mean(softmax(x,1)^2)/2."""
    v = torch.softmax(logits, dim=1)
    return -(v ** 2).mean() / 2


def self_information_map(logits: torch.Tensor) -> torch.Tensor:
    """utils/image.py:121-124 ``entropy_map``: -p*log2(p+1e-30)/log2(C), same shape."""
    v = torch.softmax(logits, dim=1)
    return -(v * torch.log2(v + 1e-30)) / math.log2(v.shape[1])


def advent_loss(y_pred: torch.Tensor, label: float) -> torch.Tensor:
    """losses/advent.py:10-18: BCE-with-logits (mean) against a constant domain label.
    (The reference builds the label with ``.to(y_pred.get_device())`` which fails on CPU
    tensors -- SURVEY 8c -- so the oracle states what :8,:16 compute.)"""
    return torch.nn.functional.binary_cross_entropy_with_logits(
        y_pred, torch.full_like(y_pred, float(label)))


def softmax_loss_with_grad(logits: torch.Tensor, kind: str, eta: Optional[float] = None,
                           grad_scale: float = 1.0):
    """fwd + autograd bwd of entropy / max-squares, as the UDA steps drive them
    (uda/entropy_minimization.py:27-32: ``loss *= w; loss.backward()``)."""
    x = logits.detach().clone().requires_grad_(True)
    loss = entropy_loss(x, eta) if kind == "entropy" else max_square_loss(x)
    (loss * grad_scale).backward()
    return loss.detach(), x.grad


def self_information_backward(logits: torch.Tensor, upstream: torch.Tensor):
    """autograd through ``entropy_map`` with a dense upstream gradient
    (uda/adversarial_entropy_minimization.py:91-92,104-110)."""
    x = logits.detach().clone().requires_grad_(True)
    out = self_information_map(x)
    out.backward(upstream)
    return out.detach(), x.grad


# --------------------------------------------------------------------------- #
# target rasteriser (the step before the path)
# --------------------------------------------------------------------------- #
def gaussian_radius(det_size, min_overlap=0.7):
    """utils/image.py:8-28: smallest of the three CornerNet roots (each divided by 2, as the reference does)."""
    import numpy as np
    height, width = det_size
    b1 = height + width
    c1 = width * height * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + np.sqrt(b1 ** 2 - 4 * 1 * c1)) / 2
    b2 = 2 * (height + width)
    c2 = (1 - min_overlap) * width * height
    r2 = (b2 + np.sqrt(b2 ** 2 - 4 * 4 * c2)) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (height + width)
    c3 = (min_overlap - 1) * width * height
    r3 = (b3 + np.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
    return min(r1, r2, r3)


def draw_gaussian(plane, cx, cy, radius):
    """utils/image.py:31-57: max-blend the (2r+1)^2 gaussian, sigma = (2r+1)/6, clipped at the borders."""
    import numpy as np
    diameter = 2 * radius + 1
    sigma = diameter / 6
    m = (diameter - 1.) / 2.
    y, x = np.ogrid[-m:m + 1, -m:m + 1]
    g = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
    g[g < np.finfo(g.dtype).eps * g.max()] = 0
    h, w = plane.shape
    left, right = min(cx, radius), min(w - cx, radius + 1)
    top, bottom = min(cy, radius), min(h - cy, radius + 1)
    view = plane[cy - top:cy + bottom, cx - left:cx + right]
    part = g[radius - top:radius + bottom, radius - left:radius + right]
    if min(part.shape) > 0 and min(view.shape) > 0:
        np.maximum(view, part, out=view)


def raster_targets(boxes, classes, n_obj, num_classes, height, width):
    """datasets/coco.py:168-215 for boxes already in heat-map pixels: numpy arrays in, the batch dict out."""
    import numpy as np
    B, M = classes.shape
    hm = np.zeros((B, num_classes, height, width), dtype=np.float32)
    wh = np.zeros((B, M, 2), dtype=np.float32)
    reg = np.zeros((B, M, 2), dtype=np.float32)
    ind = np.zeros((B, M), dtype=np.int64)
    mask = np.zeros((B, M), dtype=np.uint8)
    for b in range(B):
        for k in range(int(n_obj[b])):
            bbox = np.array(boxes[b, k], dtype=np.float64)
            bbox[[0, 2]] = np.clip(bbox[[0, 2]], 0, width - 1)          # :199-200
            bbox[[1, 3]] = np.clip(bbox[[1, 3]], 0, height - 1)
            h, w = bbox[3] - bbox[1], bbox[2] - bbox[0]
            cls = int(classes[b, k])
            if h > 0 and w > 0 and 0 <= cls < num_classes:
                radius = max(0, int(gaussian_radius((np.ceil(h), np.ceil(w)))))
                ct = np.array([(bbox[0] + bbox[2]) / 2, (bbox[1] + bbox[3]) / 2], dtype=np.float32)
                ct_int = ct.astype(np.int32)
                draw_gaussian(hm[b, cls], int(ct_int[0]), int(ct_int[1]), radius)
                wh[b, k] = 1. * w, 1. * h
                ind[b, k] = ct_int[1] * width + ct_int[0]
                reg[b, k] = ct - ct_int
                mask[b, k] = 1
    return {"hm": hm, "reg_mask": mask, "ind": ind, "wh": wh, "reg": reg}
