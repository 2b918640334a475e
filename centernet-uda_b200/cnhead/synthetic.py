"""Deterministic synthetic head tensors and targets for the five BASELINE.json configs.

Shapes and distributions follow SURVEY.md section 8(d): head-map logits ~ N(-2.19, s^2)
(the ``hm`` bias init of backends/dla.py:485), ``wh`` ~ U(0,40) (+ angle channel ~ N(0,1)),
``reg`` ~ U(0,1); targets are rasterised the way the reference dataset does it
(datasets/coco.py:191-233 with utils/image.py:8-57): one gaussian splat per object on
its class plane (exact 1.0 at the integer centre), ``ind = cy*W + cx`` (int64),
``reg_mask`` uint8, ``wh``/``reg`` rows, unused slots zero.  Everything is generated on
the CPU from ``torch.Generator().manual_seed(42 + cfg_index)`` so that the oracle and the
CUDA path see identical bits.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch


@dataclass(frozen=True)
class HeadConfig:
    name: str
    index: int                 # position in BASELINE.json:configs (seed = 42 + index)
    batch: int
    classes: int
    height: int = 128
    width: int = 128
    max_objects: int = 150     # M == K == cfg.max_detections (configs/defaults.yaml:102)
    K: int = 150
    angle: bool = False        # wh head has a third (angle) channel
    periodic: bool = False
    rotated: bool = False
    target_domain: bool = False
    objects_hi: int = 20
    advent: bool = False       # the ADVENT step (uda/adversarial_entropy_minimization.py:77-152) instead of decode

    @property
    def wh_channels(self) -> int:
        return 3 if self.angle else 2

    def bytes_per_sample(self, decode: bool = True, grad: bool = True) -> int:
        """Algorithmic bytes per sample: focal 16*C*HW (12 without grad),
        dense wh/reg gradient maps 4*(D+2)*HW, decode 4*C*HW, each UDA loss 8*C*HW."""
        hw = self.height * self.width
        if self.advent:
            # source DetectionLoss fwd+bwd, three self-information maps forward (logits in, map out: 8*C*HW each), one
            # backward through the map with a dense upstream gradient (logits + upstream in, gradient out: 12*C*HW);
            # the discriminator's [N,1,h/32,w/32] logits and their BCE are negligible
            return 16 * self.classes * hw + 4 * (self.wh_channels + 2) * hw + (3 * 8 + 12) * self.classes * hw
        n = (16 if grad else 12) * self.classes * hw
        if grad:
            n += 4 * (self.wh_channels + 2) * hw
        if decode:
            n += 4 * self.classes * hw
        if self.target_domain:
            n += 2 * 8 * self.classes * hw
        return n


CONFIGS: Dict[str, HeadConfig] = {
    "cfg1": HeadConfig("cfg1", 0, batch=1, classes=6, K=100),
    "cfg2": HeadConfig("cfg2", 1, batch=16, classes=6),
    "cfg3": HeadConfig("cfg3", 2, batch=16, classes=6, angle=True, periodic=True, rotated=True),
    "cfg4": HeadConfig("cfg4", 3, batch=16, classes=6, target_domain=True),
    "cfg5": HeadConfig("cfg5", 4, batch=128, classes=80, objects_hi=60),
    # not a BASELINE config: the kernels of one ADVENT step around the (out-of-scope) discriminator network
    "advent": HeadConfig("advent", 5, batch=16, classes=6, target_domain=True, advent=True),
}


# --------------------------------------------------------------------------- #
# target rasteriser (the step before the hot path)
# --------------------------------------------------------------------------- #
def splat_radius(box_h: float, box_w: float, min_overlap: float = 0.7) -> float:
    """Largest centre displacement keeping IoU >= min_overlap (CornerNet rule used by
    utils/image.py:8-28): the smallest root of three quadratics, one per overlap case."""
    s, area = box_h + box_w, box_h * box_w
    cases = (
        (1.0, s, area * (1 - min_overlap) / (1 + min_overlap)),        # both corners inside
        (4.0, 2 * s, (1 - min_overlap) * area),                        # both outside
        (4 * min_overlap, -2 * min_overlap * s, (min_overlap - 1) * area),
    )
    # NB: the reference divides every root by 2 (not by 2a); reproduce, don't fix.
    return min((b + np.sqrt(b * b - 4 * a * c)) / 2 for a, b, c in cases)


def splat_gaussian(plane: np.ndarray, cx: int, cy: int, radius: int) -> None:
    """Max-blend a (2r+1)^2 gaussian with sigma = (2r+1)/6 centred on (cx, cy) into a
    [H,W] plane, clipped at the borders (utils/image.py:31-57)."""
    h, w = plane.shape
    sigma = (2 * radius + 1) / 6.0
    l, r = min(cx, radius), min(w - cx, radius + 1)
    t, b = min(cy, radius), min(h - cy, radius + 1)
    if l + r <= 0 or t + b <= 0:
        return
    ys = np.arange(-t, b, dtype=np.float64)[:, None]
    xs = np.arange(-l, r, dtype=np.float64)[None, :]
    g = np.exp(-(xs * xs + ys * ys) / (2 * sigma * sigma))
    g[g < np.finfo(np.float64).eps * 1.0] = 0                       # peak value is 1
    view = plane[cy - t:cy + b, cx - l:cx + r]
    np.maximum(view, g.astype(plane.dtype), out=view)


# --------------------------------------------------------------------------- #
def make_inputs(cfg: HeadConfig, batch: Optional[int] = None, hm_sigma: float = 1.0,
                seed_offset: int = 0, sample_offset: int = 0) -> Dict[str, Dict[str, torch.Tensor]]:
    """Returns {'output': head maps, 'batch': targets[, 'target': target-domain maps]} on CPU.

    ``sample_offset`` makes a batch *slice* reproducible: samples are generated one by one
    from per-sample seeds, so rank r of G can build exactly samples [r*B/G, (r+1)*B/G) of
    the global batch without materialising the rest (cfg5 is 671 MB of ``hm`` alone)."""
    B = cfg.batch if batch is None else batch
    C, H, W, M, D = cfg.classes, cfg.height, cfg.width, cfg.max_objects, cfg.wh_channels
    out = {"hm": torch.empty(B, C, H, W), "wh": torch.empty(B, D, H, W), "reg": torch.empty(B, 2, H, W)}
    tgt = {"hm": torch.zeros(B, C, H, W), "reg_mask": torch.zeros(B, M, dtype=torch.uint8),
           "ind": torch.zeros(B, M, dtype=torch.int64), "wh": torch.zeros(B, M, D),
           "reg": torch.zeros(B, M, 2)}
    tdom = {"hm": torch.empty(B, C, H, W)} if cfg.target_domain else None
    base_seed = 42 + cfg.index + 1000 * seed_offset
    for i in range(B):
        g = torch.Generator().manual_seed(base_seed * 100003 + (sample_offset + i))
        out["hm"][i] = torch.randn(C, H, W, generator=g) * hm_sigma - 2.19
        out["wh"][i, :2] = torch.rand(2, H, W, generator=g) * 40.0
        if D == 3:
            out["wh"][i, 2] = torch.randn(H, W, generator=g)
        out["reg"][i] = torch.rand(2, H, W, generator=g)
        n_obj = int(torch.randint(1, cfg.objects_hi + 1, (1,), generator=g))
        n_obj = min(n_obj, M)
        cx = torch.randint(0, W, (n_obj,), generator=g).numpy()
        cy = torch.randint(0, H, (n_obj,), generator=g).numpy()
        cls = torch.randint(0, C, (n_obj,), generator=g).numpy()
        bw = (torch.rand(n_obj, generator=g) * 56.0 + 4.0).numpy()
        bh = (torch.rand(n_obj, generator=g) * 56.0 + 4.0).numpy()
        off = torch.rand(n_obj, 2, generator=g)
        ang = torch.rand(n_obj, generator=g) * 180.0 - 90.0
        planes = tgt["hm"][i].numpy()
        for k in range(n_obj):
            rad = max(0, int(splat_radius(np.ceil(bh[k]), np.ceil(bw[k]))))
            splat_gaussian(planes[cls[k]], int(cx[k]), int(cy[k]), rad)
            tgt["ind"][i, k] = int(cy[k]) * W + int(cx[k])
            tgt["reg_mask"][i, k] = 1
            tgt["wh"][i, k, 0], tgt["wh"][i, k, 1] = float(bw[k]), float(bh[k])
            if D == 3:
                tgt["wh"][i, k, 2] = ang[k]
            tgt["reg"][i, k] = off[k]
        if tdom is not None:
            tdom["hm"][i] = torch.randn(C, H, W, generator=g) * 1.5
    res = {"output": out, "batch": tgt}
    if tdom is not None:
        res["target"] = tdom
    return res


def loss_kwargs(cfg: HeadConfig) -> Dict[str, object]:
    """DetectionLoss ctor kwargs: the reference defaults (configs/defaults.yaml:21-26)."""
    return dict(hm_weight=1.0, wh_weight=0.1, off_weight=1.0, angle_weight=1.0,
                periodic=cfg.periodic)
