"""``losses.entropy.EntropyLoss`` (reference losses/entropy.py:5-28): Shannon entropy of the channel
softmax of raw target-domain logits, forward + gradient in one sm_100a launch (csrc/softmax_stat.cu).
The returned 0-dim loss accepts the in-place ``loss *= weight`` the UDA steps apply
(uda/entropy_minimization.py:28)."""
import torch

from cnhead import _lib as _L
from cnhead import functional as _F


class EntropyLoss(torch.nn.Module):
    def __init__(self, eta=None):
        super().__init__()
        self.eta = eta

    def forward(self, outputs, batch):
        mode = _L.SOFTMAX_ENTROPY if self.eta is None else _L.SOFTMAX_ENTROPY_ETA
        entropy_loss = _F.softmax_loss(outputs['hm'], mode, self.eta)
        return entropy_loss, {'entropy_loss': entropy_loss}
