"""Drop-in for the reference's loss.py:19-27 (BPRLoss) on the sm_100a kernels (yr_bpr_loss_fwd/_bwd)."""
import torch.nn as nn

from . import ops


class BPRLoss(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, positive_preds, negative_preds):
        return ops.bpr_loss(positive_preds, negative_preds)
