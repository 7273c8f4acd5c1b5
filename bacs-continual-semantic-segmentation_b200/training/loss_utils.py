"""Loss modules of the reference's training/loss_utils.py that sit on the BACS path
(lines 447-588), as thin nn.Modules over the fused pixel kernel:

  WeightedCrossEntropy                 training/loss_utils.py:523-588
  UnbiasedCrossEntropy                 training/loss_utils.py:492-520
  UnbiasedKnowledgeDistillationLoss    training/loss_utils.py:447-489

Same constructor arguments and call signatures.  The PLOP / iCaRL helpers of that file are
out of scope."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _cabi, ops
from ..autograd import PixelLossFunction


def _seen_cfg(seen_not_seen_probs):
    """Accepts the reference's [B,T,H,W] probability tensor or a SeenMap handle."""
    if hasattr(seen_not_seen_probs, "z"):
        return {"z": seen_not_seen_probs.z, "seen_scale": seen_not_seen_probs.scale}
    # max over heads is the only thing the loss uses (loss_utils.py:545)
    return {"seen_max": seen_not_seen_probs.detach().max(1)[0].float().contiguous()}


class WeightedCrossEntropy(nn.Module):
    """Background-aware unbiased CE: focal modulation (1 - s)^gamma of the background term
    by the seen probability, unbiased new-vs-old term, mean over ALL pixels (Q6)."""

    def __init__(self, gamma=2, old_cl=None, threshold=0.5, ignore_index=255, ukd=True):
        super().__init__()
        self.ignore_index = ignore_index
        self.old_cl = old_cl
        self.eps = 1e-4
        self.gamma = gamma
        self.base_loss = None
        self.threshold = threshold
        self.ukd = ukd

    def _custom_unbiased(self, inputs, targets, seen_not_seen_probs, task_num):
        cfg = {"mode": _cabi.PIX_WEIGHTED_CE, "old_cl": int(self.old_cl), "ukd": bool(self.ukd),
               "gamma": float(self.gamma), "threshold": float(self.threshold), "ignore_index": self.ignore_index}
        cfg.update(_seen_cfg(seen_not_seen_probs))
        loss, _, _ = PixelLossFunction.apply(inputs, None, None, None, targets, cfg)
        return loss

    def forward(self, inputs, targets, seen_not_seen_probs, task_num):
        return self._custom_unbiased(inputs, targets, seen_not_seen_probs, task_num)


class UnbiasedCrossEntropy(nn.Module):
    """MiB unbiased CE: labels below old_cl collapse onto p(old) = sum_{k<old_cl} p_k."""

    def __init__(self, old_cl=None, reduction="mean", ignore_index=255):
        super().__init__()
        if reduction != "mean":
            raise NotImplementedError("UnbiasedCrossEntropy: only reduction='mean' is on the BACS path")
        self.reduction = reduction
        self.ignore_index = ignore_index
        self.old_cl = old_cl

    def forward(self, inputs, targets):
        cfg = {"mode": _cabi.PIX_UNBIASED_CE, "old_cl": int(self.old_cl), "ignore_index": self.ignore_index}
        loss, _, _ = PixelLossFunction.apply(inputs, None, None, None, targets, cfg)
        return loss


class UnbiasedKnowledgeDistillationLoss(nn.Module):
    """MiB unbiased KD.  Instantiated on every loss object (base_loss.py:78) but only
    called by MiB / SDR, never by BACS; evaluated by the dedicated kernel
    bacs_unbiased_kd (two logit tensors per pixel)."""

    def __init__(self, reduction="mean", alpha=1.0):
        super().__init__()
        self.reduction = reduction
        self.alpha = alpha

    def forward(self, inputs, targets, mask=None):
        from ..autograd import UnbiasedKDFunction
        if self.reduction != "mean":
            raise NotImplementedError("UnbiasedKnowledgeDistillationLoss: only reduction='mean' is implemented")
        return UnbiasedKDFunction.apply(inputs, targets, mask, float(self.alpha))
