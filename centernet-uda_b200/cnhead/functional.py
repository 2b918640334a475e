"""Autograd-facing wrappers over the C ABI (one kernel launch per forward, lazy rescale on backward).

Design notes
  * Forward launches compute the loss AND the gradient for an upstream gradient of 1.0 in the same
    pass over the head maps.  ``backward`` only launches ``cnh_scale_inplace``, which reads the real
    upstream gradient from device memory and returns immediately when it is exactly 1.0 (the
    ``loss.backward()`` case) -- no host synchronisation anywhere.
  * The gradients handed to autograd are the tensors written by the forward launch (no copy), so a
    graph may be backpropagated once; a second backward raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L

_ACCURATE = os.environ.get("CNH_ACCURATE_MATH", "0") == "1"
_NO_STASH = os.environ.get("CNH_NO_STASH", "0") == "1"


def default_flags() -> int:
    return (L.FLAG_ACCURATE_MATH if _ACCURATE else 0) | (L.FLAG_NO_STASH if _NO_STASH else 0)


class HeadSpec:
    """One masked gather-L1 head: prediction map, target rows, mask and weights."""

    def __init__(self, fmap, target, mask, weight, angle_weight=1.0, angle_mode=L.ANGLE_NONE,
                 elementwise_mask=False, pairs=None):
        self.fmap, self.target, self.mask = fmap, target, mask
        self.weight, self.angle_weight = float(weight), float(angle_weight)
        self.angle_mode, self.elementwise_mask = int(angle_mode), bool(elementwise_mask)
        self.pairs = pairs          # LIMB_* modes: int32 [P,2] keypoint index pairs on the device (angle_weight = their weight)


def _as_mask(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.uint8:
        t = t.to(torch.uint8)
    return L.require(t, "mask", torch.uint8)


def fill_detloss_args(hm, gt, ind, heads: Sequence[HeadSpec], hm_weight, prob, grads, scalars, totals,
                      norm=None, norm_out=None, flags=None, b_global=None) -> L.DetLossArgs:
    B, Cc, H, W = hm.shape
    a = L.DetLossArgs()
    a.B, a.C, a.H, a.W = B, Cc, H, W
    a.M = ind.shape[1] if ind is not None else 0
    a.n_heads = len(heads)
    a.flags = default_flags() if flags is None else flags
    a.B_global = B if b_global is None else b_global
    a.hm_logits, a.hm_gt, a.prob = hm.data_ptr(), gt.data_ptr(), prob.data_ptr()
    a.grad_hm = L.ptr(grads[0]) if grads is not None else None
    a.ind = L.ptr(ind)
    a.hm_weight = float(hm_weight)
    for i, h in enumerate(heads):
        hd = a.heads[i]
        hd.map, hd.target, hd.mask = h.fmap.data_ptr(), h.target.data_ptr(), h.mask.data_ptr()
        hd.grad = L.ptr(grads[1 + i]) if grads is not None else None
        hd.D = h.fmap.shape[1]
        hd.angle_mode = h.angle_mode
        hd.elementwise_mask = 1 if h.elementwise_mask else 0
        hd.weight, hd.angle_weight = h.weight, h.angle_weight
        pairs = getattr(h, "pairs", None)
        hd.pairs, hd.n_pairs = (pairs.data_ptr(), pairs.shape[0]) if pairs is not None else (None, 0)
    a.scalars, a.totals = L.ptr(scalars), L.ptr(totals)
    a.norm, a.norm_out = L.ptr(norm), L.ptr(norm_out)
    return a


def _check_heads(hm, gt, ind, heads):
    B, Cc, H, W = hm.shape
    if gt.shape != hm.shape:
        raise RuntimeError(f"cnhead: batch['hm'] {tuple(gt.shape)} does not match output['hm'] {tuple(hm.shape)}")
    if len(heads) > L.MAX_HEADS:
        raise RuntimeError("cnhead: at most 3 regression heads")
    M = ind.shape[1]
    if ind.shape[0] != B:
        raise RuntimeError("cnhead: batch['ind'] batch dimension mismatch")
    for h in heads:
        D = h.fmap.shape[1]
        if h.fmap.shape[0] != B or tuple(h.fmap.shape[2:]) != (H, W):
            raise RuntimeError(f"cnhead: head map {tuple(h.fmap.shape)} does not match heat map {tuple(hm.shape)}")
        if tuple(h.target.shape) != (B, M, D):
            raise RuntimeError(f"cnhead: head target {tuple(h.target.shape)} != {(B, M, D)}")
        want = (B, M, D) if h.elementwise_mask else (B, M)
        if tuple(h.mask.shape) != want:
            raise RuntimeError(f"cnhead: head mask {tuple(h.mask.shape)} != {want}")


class Candidates:
    """Peak candidates a detection-loss launch emitted for its own probability map (include/cnhead.h: cnh_cand).
    ``decode`` (below) uses them when it is handed that very map: the heat map is then not read again."""

    _dirty = {}                                         # workspace key -> Candidates whose lists were never decoded

    def __init__(self, K: int, device):
        self.c = L.Cand()
        self.K = int(K)
        nbytes = 0
        self.ws = None
        self.key = None
        self.prob_ptr, self.prob_version, self.stream = None, None, None
        self.device = device

    def prepare(self, B: int):
        """workspace of this (device, stream, batch): zeroed once; cleaned again here if its last lists were dropped"""
        nbytes = L.lib().cnh_cand_workspace_bytes(B)
        kind = f"cand:{B}"
        self.ws = L.workspace(kind, nbytes, self.device)
        self.key = (self.ws.data_ptr(), kind)
        stale = Candidates._dirty.pop(self.key, None)
        if stale is not None and stale.c.G > 0:
            n = L.lib().cnh_cand_state_bytes(C.byref(stale.c))
            self.ws[:n].zero_()
        self.c.workspace, self.c.workspace_bytes = self.ws.data_ptr(), self.ws.numel()
        self.c.K, self.c.G = self.K, 0

    def emitted(self, prob: torch.Tensor):
        if self.c.G > 0:
            self.prob_ptr, self.prob_version, self.stream = prob.data_ptr(), prob._version, L.stream_ptr()
            Candidates._dirty[self.key] = self

    def usable_for(self, heat: torch.Tensor, K: int) -> bool:
        return (self.c.G > 0 and self.prob_ptr == heat.data_ptr() and self.prob_version == heat._version and
                K <= self.K and tuple(heat.shape) == (self.c.B, self.c.C, self.c.H, self.c.W) and
                self.stream == L.stream_ptr() and Candidates._dirty.get(self.key) is self)

    def consumed(self):
        Candidates._dirty.pop(self.key, None)
        self.c.G = 0


class _DetectionLossFn(torch.autograd.Function):
    """inputs: hm logits, then one map per head.  Outputs: scalars[8], clamped prob, totals[24]."""

    @staticmethod
    def forward(ctx, meta, hm, *maps):
        gt, ind, specs, hm_weight, cand = meta
        heads = [HeadSpec(m, s.target, s.mask, s.weight, s.angle_weight, s.angle_mode, s.elementwise_mask, s.pairs)
                 for m, s in zip(maps, specs)]
        need_grad = any(ctx.needs_input_grad[1:])
        prob = torch.empty_like(hm)
        grads = [torch.empty_like(hm)] + [torch.empty_like(m) for m in maps] if need_grad else None
        scalars = torch.empty(L.SCALARS, dtype=torch.float32, device=hm.device)
        totals = torch.empty(L.TOTALS, dtype=torch.int64, device=hm.device)
        a = fill_detloss_args(hm, gt, ind, heads, hm_weight, prob, grads, scalars, totals)
        if cand is not None:
            cand.prepare(hm.shape[0])
            a.cand = C.pointer(cand.c)
        nbytes = L.lib().cnh_detloss_workspace_bytes(C.byref(a))
        ws = L.workspace("detloss", nbytes, hm.device)
        L.check(L.lib().cnh_detloss_fused(C.byref(a), ws.data_ptr(), ws.numel(), L.stream_ptr()), "detloss_fused")
        if cand is not None:
            cand.emitted(prob)
        ctx.grads = grads
        ctx.used = False
        ctx.mark_non_differentiable(prob, totals)
        return scalars, prob, totals

    @staticmethod
    def backward(ctx, g_scalars, _g_prob, _g_totals):
        if ctx.used:
            raise RuntimeError("cnhead: the fused DetectionLoss graph can be backpropagated only once "
                               "(its gradients were produced by the forward launch); run forward again")
        if ctx.grads is None:
            raise RuntimeError("cnhead: DetectionLoss was run without gradients")
        ctx.used = True
        grads, ctx.grads = ctx.grads, None      # hand over the only reference: AccumulateGrad then keeps the
        g = L.require(g_scalars, "grad_output")  # tensors as .grad instead of cloning them (3 copy launches)
        s = L.ScaleArgs()
        s.n_tensors = len(grads)
        base = g.data_ptr()
        for i, t in enumerate(grads):
            s.data[i] = t.data_ptr()
            s.count[i] = t.numel()
            s.fa[i] = base                      # d/d(total loss)
            s.fb[i] = base + 4 * (1 + i)        # d/d(hm_loss | head-i loss) if a stat is differentiated
        L.check(L.lib().cnh_scale_inplace(C.byref(s), L.stream_ptr()), "scale_inplace")
        needs = ctx.needs_input_grad[1:]
        out = [None] + [gr if need else None for gr, need in zip(grads, needs)]
        del grads
        return tuple(out)


class _Spec:
    __slots__ = ("target", "mask", "weight", "angle_weight", "angle_mode", "elementwise_mask", "pairs")


def _spec_of(h: "HeadSpec") -> "_Spec":
    sp = _Spec()
    sp.target = L.require(h.target, "head target")
    sp.mask = _as_mask(h.mask)
    sp.weight, sp.angle_weight = h.weight, h.angle_weight
    sp.angle_mode, sp.elementwise_mask = h.angle_mode, h.elementwise_mask
    sp.pairs = None
    if h.angle_mode in (L.LIMB_SQRT, L.LIMB_L1):
        if h.pairs is None:
            raise RuntimeError("cnhead: a limb-length head needs its keypoint index pairs")
        sp.pairs = L.require(h.pairs, "kp_indices", torch.int32)
        if sp.pairs.dim() != 2 or sp.pairs.shape[1] != 2:
            raise RuntimeError(f"cnhead: kp_indices must be [P,2], got {tuple(sp.pairs.shape)}")
    return sp


def detection_loss(hm: torch.Tensor, gt: torch.Tensor, ind: torch.Tensor, heads: Sequence[HeadSpec],
                   hm_weight: float = 1.0, decode_K: Optional[int] = None):
    """Fused DetectionLoss core.  Returns (scalars[8], prob, totals[24] int64); see include/cnhead.h.
    ``decode_K``: also emit the peak candidates of the probability map for a later ``decode(prob, ..., K <= decode_K)``
    (large heat maps only; the returned map then carries them as ``prob._cnh_cand``)."""
    hm = L.require(hm, "output['hm']")
    gt = L.require(gt, "batch['hm']")
    ind = L.require(ind, "batch['ind']", torch.int64)
    specs, maps = [], []
    for h in heads:
        specs.append(_spec_of(h))
        maps.append(L.require(h.fmap, "head map"))
    _check_heads(hm, gt, ind, [HeadSpec(m, s.target, s.mask, s.weight, s.angle_weight, s.angle_mode,
                                        s.elementwise_mask, s.pairs) for m, s in zip(maps, specs)])
    cand = Candidates(decode_K, hm.device) if decode_K else None
    scalars, prob, totals = _DetectionLossFn.apply((gt, ind, specs, float(hm_weight), cand), hm, *maps)
    if cand is not None and cand.c.G > 0:
        prob._cnh_cand = cand
    return scalars, prob, totals


# ----------------------------------------------------------------------------------------------
# channel-softmax losses
# ----------------------------------------------------------------------------------------------
class _SoftmaxLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mode, eta, n_total):
        N, Cc, H, W = x.shape
        need_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(x) if need_grad else None
        out = torch.empty((), dtype=torch.float32, device=x.device)
        nbytes = L.lib().cnh_softmax_workspace_bytes(N, Cc, H, W)
        ws = L.workspace("softmax", nbytes, x.device)
        L.check(L.lib().cnh_softmax_loss(x.data_ptr(), L.ptr(grad), out.data_ptr(), N, Cc, H, W,
                                         n_total if n_total else N, mode, float(eta if eta is not None else 0.0),
                                         ws.data_ptr(), ws.numel(), L.stream_ptr()), "softmax_loss")
        ctx.grad = grad
        ctx.used = False
        return out

    @staticmethod
    def backward(ctx, g_out):
        if ctx.used:
            raise RuntimeError("cnhead: this fused loss can be backpropagated only once; run forward again")
        ctx.used = True
        g = L.require(g_out, "grad_output")
        s = L.ScaleArgs()
        s.n_tensors = 1
        s.data[0], s.count[0], s.fa[0], s.fb[0] = ctx.grad.data_ptr(), ctx.grad.numel(), g.data_ptr(), None
        L.check(L.lib().cnh_scale_inplace(C.byref(s), L.stream_ptr()), "scale_inplace")
        grad, ctx.grad = ctx.grad, None         # sole reference -> no clone in AccumulateGrad
        return grad, None, None, None


def softmax_loss(x: torch.Tensor, mode: int, eta: Optional[float] = None, n_total: int = 0) -> torch.Tensor:
    x = L.require(x, "outputs['hm']")
    if x.dim() != 4:
        raise RuntimeError(f"cnhead: expected [N,C,H,W] logits, got {tuple(x.shape)}")
    return _SoftmaxLossFn.apply(x, mode, eta, n_total)


class _EntropyMapFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        N, Cc, H, W = x.shape
        out = torch.empty_like(x)
        L.check(L.lib().cnh_entropy_map_fwd(x.data_ptr(), out.data_ptr(), N, Cc, H, W, L.stream_ptr()),
                "entropy_map_fwd")
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = L.require(g, "grad_output")
        N, Cc, H, W = x.shape
        gin = torch.empty_like(x)
        L.check(L.lib().cnh_entropy_map_bwd(x.data_ptr(), g.data_ptr(), gin.data_ptr(), N, Cc, H, W,
                                            L.stream_ptr()), "entropy_map_bwd")
        return gin


def entropy_map(x: torch.Tensor) -> torch.Tensor:
    x = L.require(x, "hm")
    if x.dim() != 4:
        raise RuntimeError(f"cnhead: expected [N,C,H,W] logits, got {tuple(x.shape)}")
    return _EntropyMapFn.apply(x)


class _BceConstFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, label):
        need_grad = ctx.needs_input_grad[0]
        grad = torch.empty_like(y) if need_grad else None
        out = torch.empty((), dtype=torch.float32, device=y.device)
        L.check(L.lib().cnh_bce_const(y.data_ptr(), L.ptr(grad), out.data_ptr(), y.numel(), float(label),
                                      L.stream_ptr()), "bce_const")
        ctx.grad = grad
        return out

    @staticmethod
    def backward(ctx, g_out):
        return ctx.grad * g_out, None          # a [N,1,h/32,w/32] discriminator map: tiny


def bce_const(y: torch.Tensor, label: float) -> torch.Tensor:
    return _BceConstFn.apply(L.require(y, "y_pred"), float(label))


# ----------------------------------------------------------------------------------------------
# decode
# ----------------------------------------------------------------------------------------------
def decode(heat, wh, reg=None, kps=None, K=100, rotated=False, apply_sigmoid=False, box_scale=1.0,
           return_inds=False, score_threshold=None):
    """backends/decode.py:35-76 in one launch.  Extras over the reference signature (its callers' epilogues,
    SURVEY 8f N1/N4): ``apply_sigmoid`` (export.py:31-33), ``box_scale`` (uda/base.py:90 ``down_ratio``),
    ``return_inds`` (flat peak indices) and ``score_threshold`` -> an extra int32 ``[B]`` tensor with the number
    of rows per sample whose score reaches it (rows are sorted, so they are the first ones;
    evaluation/coco.py:266-267).  Returns dets[, kps][, inds][, counts]."""
    heat_in = heat
    heat = L.require(heat.detach(), "heat")
    wh = L.require(wh.detach(), "wh")
    reg = L.require(reg.detach(), "reg") if reg is not None else None
    kps = L.require(kps.detach(), "kps") if kps is not None else None
    if heat.dim() != 4 or wh.dim() != 4:
        raise RuntimeError("cnhead: decode expects [B,C,H,W] heat and [B,D,H,W] wh")
    B, Cc, H, W = heat.shape
    if wh.shape[0] != B or tuple(wh.shape[2:]) != (H, W):
        raise RuntimeError(f"cnhead: wh {tuple(wh.shape)} does not match heat {tuple(heat.shape)}")
    if reg is not None and tuple(reg.shape) != (B, 2, H, W):
        raise RuntimeError(f"cnhead: reg {tuple(reg.shape)} != {(B, 2, H, W)}")
    K = int(K)
    a = L.DecodeArgs()
    a.B, a.C, a.H, a.W, a.K, a.D = B, Cc, H, W, K, wh.shape[1]
    a.rotated = 1 if rotated else 0
    a.nk = kps.shape[1] // 2 if kps is not None else 0
    dets = torch.empty(B, K, 7 if rotated else 6, dtype=torch.float32, device=heat.device)
    inds = torch.empty(B, K, dtype=torch.int64, device=heat.device) if return_inds else None
    kout = torch.empty(B, K, a.nk, 2, dtype=torch.float32, device=heat.device) if kps is not None else None
    a.heat, a.wh, a.reg, a.kps = heat.data_ptr(), wh.data_ptr(), L.ptr(reg), L.ptr(kps)
    a.dets, a.inds_out, a.kps_out = dets.data_ptr(), L.ptr(inds), L.ptr(kout)
    a.apply_sigmoid = 1 if apply_sigmoid else 0
    a.box_scale = float(box_scale)
    counts = None
    if score_threshold is not None:
        counts = torch.empty(B, dtype=torch.int32, device=heat.device)
        a.counts_out, a.score_threshold = counts.data_ptr(), float(score_threshold)
    cand = getattr(heat_in, "_cnh_cand", None)
    if cand is not None and not apply_sigmoid and cand.usable_for(heat, K):
        # the loss launch that wrote this very map left its peak candidates: the heat map is not read again
        L.check(L.lib().cnh_decode_candidates(C.byref(a), C.byref(cand.c), L.stream_ptr()), "decode_candidates")
        cand.consumed()
    else:
        nbytes = L.lib().cnh_decode_workspace_bytes(C.byref(a))
        if nbytes == 0:
            L.check(-2, "decode")
        ws = L.workspace(f"decode:{B}x{Cc}x{H}x{W}:{K}", nbytes, heat.device)   # layout depends on the dims
        L.check(L.lib().cnh_decode(C.byref(a), ws.data_ptr(), ws.numel(), L.stream_ptr()), "decode")
    res = (dets,)
    if kps is not None:
        res += (kout,)
    if return_inds:
        res += (inds,)
    if counts is not None:
        res += (counts,)
    return res[0] if len(res) == 1 else res


# ----------------------------------------------------------------------------------------------
# target rasteriser (the step before the loss)
# ----------------------------------------------------------------------------------------------
def raster_targets(boxes, classes, n_obj, num_classes, height, width) -> Dict[str, torch.Tensor]:
    """datasets/coco.py:168-215 on the device: from ``boxes [B,M,4]`` fp32 (x1,y1,x2,y2 in heat-map pixels),
    ``classes [B,M]`` int32 and ``n_obj [B]`` int32 to the batch dict DetectionLoss consumes --
    ``{'hm' [B,C,H,W], 'reg_mask' [B,M] u8, 'ind' [B,M] i64, 'wh' [B,M,2], 'reg' [B,M,2]}``.
    The host then ships B*M*20 bytes per step instead of the 4*B*C*H*W-byte heat-map target."""
    boxes = L.require(boxes, "boxes")
    classes = L.require(classes, "classes", torch.int32)
    n_obj = L.require(n_obj, "n_obj", torch.int32)
    if boxes.dim() != 3 or boxes.shape[2] != 4:
        raise RuntimeError(f"cnhead: boxes must be [B,M,4], got {tuple(boxes.shape)}")
    B, M = boxes.shape[:2]
    if tuple(classes.shape) != (B, M) or tuple(n_obj.shape) != (B,):
        raise RuntimeError("cnhead: classes must be [B,M] and n_obj [B]")
    dev = boxes.device
    out = {"hm": torch.empty(B, num_classes, height, width, dtype=torch.float32, device=dev),
           "reg_mask": torch.empty(B, M, dtype=torch.uint8, device=dev),
           "ind": torch.empty(B, M, dtype=torch.int64, device=dev),
           "wh": torch.empty(B, M, 2, dtype=torch.float32, device=dev),
           "reg": torch.empty(B, M, 2, dtype=torch.float32, device=dev)}
    a = L.RasterArgs()
    a.B, a.C, a.H, a.W, a.M = B, int(num_classes), int(height), int(width), M
    a.min_overlap_num, a.min_overlap_den = 7, 10            # utils/image.py:8 default, the only value callers use
    a.boxes, a.classes, a.n_obj = boxes.data_ptr(), classes.data_ptr(), n_obj.data_ptr()
    a.hm, a.wh, a.reg = out["hm"].data_ptr(), out["wh"].data_ptr(), out["reg"].data_ptr()
    a.ind, a.reg_mask = out["ind"].data_ptr(), out["reg_mask"].data_ptr()
    L.check(L.lib().cnh_raster_targets(C.byref(a), L.stream_ptr()), "raster_targets")
    return out
