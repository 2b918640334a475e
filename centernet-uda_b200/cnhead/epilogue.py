"""Host epilogue of the detection decode: what ``Model.get_detections`` (uda/base.py:73-94) and
``export.CenterNet.forward`` (export.py:27-56) do after ``decode_detection`` -- scale the box columns by
``down_ratio``, move the detections to the host, split them into boxes / scores / classes and (in the
evaluator, evaluation/coco.py:266-267) drop rows under the score threshold -- with the scale and the
threshold count folded into the decode launch and the device->host copy made asynchronous into pinned
memory (the reference blocks in ``.cpu().numpy()``).

    fetch = DetectionsFetcher(max_detections=150, down_ratio=4, score_threshold=0.3)
    pending = fetch.launch(out['hm'], out['wh'], out['reg'])      # returns at once; kernels + copies queued
    ...                                                           # more GPU work may be queued meanwhile
    res = pending.result()                                        # waits on the copy's event only
    res['pred_boxes'][i][:res['counts'][i]]                       # detections of sample i above the threshold
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import functional as F


class PendingDetections:
    def __init__(self, event, dets, kps, counts, rotated):
        self._event, self._dets, self._kps, self._counts, self._rotated = event, dets, kps, counts, rotated

    def done(self) -> bool:
        return self._event.query()

    def result(self) -> Dict[str, np.ndarray]:
        """Same keys as uda/base.py:122-131 for the predictions (+ 'counts' when a threshold was given)."""
        self._event.synchronize()
        dets = self._dets.numpy()
        box_idx = 5 if self._rotated else 4
        out = {'pred_boxes': dets[:, :, :box_idx], 'pred_scores': dets[:, :, box_idx],
               'pred_classes': dets[:, :, box_idx + 1].astype(np.int32)}
        if self._kps is not None:
            out['pred_kps'] = self._kps.numpy()
        if self._counts is not None:
            out['counts'] = self._counts.numpy()
        return out


class DetectionsFetcher:
    """Reusable pinned staging (``depth`` result slots, round-robin) for decode -> host."""

    def __init__(self, max_detections: int, down_ratio: float = 4.0, rotated: bool = False,
                 score_threshold: Optional[float] = None, apply_sigmoid: bool = False, depth: int = 2):
        self.K, self.down_ratio, self.rotated = int(max_detections), float(down_ratio), bool(rotated)
        self.score_threshold, self.apply_sigmoid = score_threshold, bool(apply_sigmoid)
        self.depth, self._slots, self._next = max(1, int(depth)), {}, 0
        self._pending = {}                     # slot key -> event of the copy that last targeted its pinned buffers

    def _slot(self, B, nk):
        key = (self._next % self.depth, B, nk)
        self._next += 1
        prev = self._pending.get(key)
        if prev is not None:                   # a result of `depth` launches ago may still be in flight into these
            prev.synchronize()                 # buffers (or unread): never overwrite under a pending copy
        self._last_key = key
        if key not in self._slots:
            self._slots[key] = (torch.empty(B, self.K, 7 if self.rotated else 6).pin_memory(),
                                torch.empty(B, self.K, nk, 2).pin_memory() if nk else None,
                                torch.empty(B, dtype=torch.int32).pin_memory())
        return self._slots[key]

    def launch(self, heat, wh, reg=None, kps=None) -> PendingDetections:
        res = F.decode(heat, wh, reg, kps, K=self.K, rotated=self.rotated, apply_sigmoid=self.apply_sigmoid,
                       box_scale=self.down_ratio, score_threshold=self.score_threshold)
        res = res if isinstance(res, tuple) else (res,)
        dets = res[0]
        kout = res[1] if kps is not None else None
        counts = res[-1] if self.score_threshold is not None else None
        h_dets, h_kps, h_counts = self._slot(dets.shape[0], kps.shape[1] // 2 if kps is not None else 0)
        h_dets.copy_(dets, non_blocking=True)
        if kout is not None:
            h_kps.copy_(kout, non_blocking=True)
        if counts is not None:
            h_counts.copy_(counts, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dets.device))
        self._pending[self._last_key] = ev
        return PendingDetections(ev, h_dets, h_kps if kout is not None else None,
                                 h_counts if counts is not None else None, self.rotated)
