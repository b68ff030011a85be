"""Losses of the hot path — reference utils/loss.py:17-24 (`sdf_bce_loss`).

Class-surface version (takes a `pred` tensor); the fused kernel computes the same expression per point in
registers (`csrc/shine_b200.cu`, "sdf_bce_loss" block)."""
import torch
import torch.nn.functional as F


def sdf_bce_loss(pred, label, sigma, weight, weighted=False, bce_reduction="mean"):
    """BCE-with-logits between the predicted logits and the occupancy target sigmoid(label / sigma)."""
    target = torch.sigmoid(label / sigma)
    return F.binary_cross_entropy_with_logits(pred, target, weight=weight if weighted else None,
                                              reduction=bce_reduction)
