"""``backends.decode.decode_detection`` -- drop-in for backends/decode.py:35-76 of the reference,
backed by ONE sm_100a launch (csrc/decode.cu): TMA-staged 3x3 peak NMS, radix-select top-K over
C*H*W per sample, gather of reg/wh/angle/keypoints straight from NCHW, box assembly.

Same signature and output: ``[B,K,6]`` = (x1,y1,x2,y2,score,class) or, ``rotated``, ``[B,K,7]`` =
(x,y,w,h,angle,score,class), sorted by score descending; with ``kps`` a tuple ``(dets, kps[B,K,nk,2])``.
Ties are broken by the LOWER flat index ``c*H*W + y*W + x`` (torch.topk leaves them unspecified).
``heat`` must hold probabilities in [0,1] (the reference relies on DetectionLoss / export having applied
``clamp(sigmoid)``, backends/decode.py:39).  Only the 3x3 NMS window the reference ever uses is
implemented."""
from cnhead import functional as _F
from cnhead._dropin import reexport as _reexport


def decode_detection(heat, wh, reg=None, kps=None, K=100, rotated=False, nms_size=3):
    if nms_size != 3:
        raise NotImplementedError(f"decode_detection: nms_size={nms_size}; only the 3x3 window is implemented")
    return _F.decode(heat, wh, reg, kps, K=K, rotated=rotated)


_reexport(__name__, __file__, globals())
