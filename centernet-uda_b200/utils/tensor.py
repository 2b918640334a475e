"""``utils.tensor`` helpers of the reference (utils/tensor.py:5-25), kept importable for callers that
use them directly.  The hot path does not call them: sigmoid+clamp is fused into the loss / decode
kernels and the gathers read NCHW in place (no transpose copy)."""
import torch


def _sigmoid(x):
    """clamp(sigmoid_(x), 1e-4, 1-1e-4): mutates ``x`` to the unclamped sigmoid like the reference."""
    return torch.clamp(x.sigmoid_(), min=1e-4, max=1 - 1e-4)


def _gather_feat(feat, ind, mask=None):
    dim = feat.size(2)
    feat = feat.gather(1, ind.unsqueeze(2).expand(ind.size(0), ind.size(1), dim))
    if mask is not None:
        feat = feat[mask.unsqueeze(2).expand_as(feat)].view(-1, dim)
    return feat


def _transpose_and_gather_feat(feat, ind):
    """[B,D,H,W], [B,M] -> [B,M,D] without materialising the NHWC transpose."""
    b, d = feat.shape[:2]
    return feat.reshape(b, d, -1).gather(2, ind.unsqueeze(1).expand(-1, d, -1)).transpose(1, 2).contiguous()
