"""``losses.advent.AdventLoss`` (reference losses/advent.py:5-18): mean BCE-with-logits of the
discriminator output against a constant domain label.  The reference builds the label tensor on the
CPU and copies it to the device every call; here the label is a kernel argument (csrc/softmax_stat.cu)."""
import torch

from cnhead import functional as _F


class AdventLoss(torch.nn.Module):
    def forward(self, y_pred, y_true):
        advent_loss = _F.bce_const(y_pred, float(y_true))
        return advent_loss, {'advent_loss': advent_loss}
