"""``losses.max_square.MaxSquareLoss`` (reference losses/max_square.py:5-14): -mean(softmax(x,1)^2)/2,
forward + gradient in one sm_100a launch (csrc/softmax_stat.cu)."""
import torch

from cnhead import _lib as _L
from cnhead import functional as _F


class MaxSquareLoss(torch.nn.Module):
    def forward(self, outputs, batch):
        loss = _F.softmax_loss(outputs['hm'], _L.SOFTMAX_MAX_SQUARE)
        return loss, {'max_square_loss': loss}
