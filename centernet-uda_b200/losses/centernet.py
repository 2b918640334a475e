"""``losses.centernet.DetectionLoss`` -- drop-in for the reference module of the same dotted name
(losses/centernet.py:7-56), backed by ONE fused sm_100a launch (csrc/detloss.cu) instead of ~100
eager kernels.  Same constructor kwargs (configs/defaults.yaml:21-26), same call
``crit(output_dict, batch_dict) -> (loss, stats)``, same stat keys, same post-condition
``output['hm'] = clamp(sigmoid(output['hm']), 1e-4, 1-1e-4)`` that decode relies on
(uda/base.py:76-77).

Deliberate differences (no caller depends on them):
  * the head tensors and ``batch['wh'|'reg'|'kps']`` are NOT modified in place (the reference
    sigmoid-s ``output['hm']`` storage and masks the targets in place);
  * the returned probability map carries no autograd history;
  * a graph can be backpropagated once (gradients are produced by the forward launch).
CUDA fp32 tensors only: CPU tensors raise ``RuntimeError`` (there is no fallback path).
"""
import os as _os

import torch

from cnhead import _lib as _L
from cnhead import functional as _F
from cnhead._dropin import reexport as _reexport


class DetectionLoss(torch.nn.Module):
    def __init__(self, hm_weight, wh_weight, off_weight, kp_weight=None, angle_weight=1.0, periodic=False,
                 kp_indices=None, kp_distance_weight=0.1, kp_distance_weight_l1=False, max_detections=None):
        """Same kwargs as the reference (configs/defaults.yaml:21-26) plus ``max_detections`` (default: the
        environment variable CNH_FUSE_DECODE_K, else off): the loss launch also emits the peak candidates of its
        probability map for a later ``decode_detection(output['hm'], ..., K <= max_detections)``, which then does not
        read the heat map again (large maps only: the streaming schedule; small ones decode as usual)."""
        super().__init__()
        if max_detections is None and _os.environ.get("CNH_FUSE_DECODE_K"):
            max_detections = int(_os.environ["CNH_FUSE_DECODE_K"])
        self.max_detections = int(max_detections) if max_detections else None
        self.hm_weight, self.wh_weight, self.off_weight = hm_weight, wh_weight, off_weight
        self.angle_weight, self.periodic = angle_weight, periodic
        self.with_keypoints = kp_weight is not None or kp_indices is not None
        self.kp_weight = kp_weight
        self.kp_indices = torch.tensor(kp_indices) if kp_indices else None
        self.kp_distance_weight, self.kp_distance_weight_l1 = kp_distance_weight, kp_distance_weight_l1

    def _heads(self, output, batch):
        wh = output['wh']
        mode = _L.ANGLE_NONE
        if wh.shape[1] == 3:                      # losses/centernet.py:112 / :15-19
            mode = _L.ANGLE_PERIODIC if self.periodic else _L.ANGLE_SIGMOID
        elif self.periodic:
            raise RuntimeError("DetectionLoss(periodic=True) needs a 3-channel 'wh' head")
        heads = [_F.HeadSpec(wh, batch['wh'], batch['reg_mask'], self.wh_weight, self.angle_weight, mode),
                 _F.HeadSpec(output['reg'], batch['reg'], batch['reg_mask'], self.off_weight)]
        if self.with_keypoints:                   # losses/centernet.py:143-151 (+ the limb-length term, :153-187)
            kps = output['kps']
            mode, pairs = _L.ANGLE_NONE, None
            if self.kp_indices is not None:
                mode = _L.LIMB_L1 if self.kp_distance_weight_l1 else _L.LIMB_SQRT
                pairs = self._pairs_on(kps.device)
            heads.append(_F.HeadSpec(kps, batch['kps'], batch['kp_reg_mask'],
                                     1.0 if self.kp_weight is None else self.kp_weight,
                                     angle_weight=self.kp_distance_weight, angle_mode=mode,
                                     elementwise_mask=True, pairs=pairs))
        return heads

    def _pairs_on(self, device):
        """the reference's kps_weight_indices as an int32 [P,2] device table (cached per device)"""
        cache = self.__dict__.setdefault('_pairs_cache', {})
        if device not in cache:
            cache[device] = self.kp_indices.to(device=device, dtype=torch.int32).contiguous()
        return cache[device]

    def forward(self, output, batch):
        heads = self._heads(output, batch)
        scalars, prob, totals = _F.detection_loss(output['hm'], batch['hm'], batch['ind'], heads,
                                                    self.hm_weight, decode_K=self.max_detections)
        output['hm'] = prob                        # losses/centernet.py:34
        loss, hm_loss, wh_loss, off_loss = scalars[0], scalars[1], scalars[2], scalars[3]
        stats = {'centernet_loss': loss, 'hm_loss': hm_loss, 'wh_loss': wh_loss, 'off_loss': off_loss}
        if self.with_keypoints:
            stats['kp_loss'] = scalars[4]           # main term + limb-length term, both from the fused launch
        self.last_totals = totals
        return loss, stats


_reexport(__name__, __file__, globals())
