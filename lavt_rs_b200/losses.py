"""Training criteria of the reference (losses.py:7-243; selected by ``--loss`` in train.py:703-713), evaluated on the autograd-connected
logits that ``model(x, text, l_mask)`` returns in training mode.  ``loss.backward()`` hands d loss / d logits to the hand-written backward
of the hot path (lavt_rs_b200/training.py::SegmentFunction); the criteria themselves are a few reductions over a (B, 2, H, W) tensor and are
not on the hot path (the default weighted cross-entropy also exists as a fused kernel: ``lavt_cross_entropy``).

    cross_entropy_loss   [0.9, 1.1]-weighted CE                                    (:7-11)
    MultiClassDiceLoss   soft Dice with squared-probability cardinality            (:38-77)
    DiceFocalLoss        Dice * dice_rate + binary focal (alpha .25, gamma 2) * focal_rate   (:80-139)
    DiceBoundaryLoss     Dice * dice_rate + BoundaryLoss * boundary_rate           (:142-188)
    BoundaryLoss         boundary F1 on max-pooled boundary maps (theta0 3, theta 5)          (:191-243)
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def cross_entropy_loss(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return F.cross_entropy(input, target, weight=torch.tensor([0.9, 1.1], device=input.device, dtype=input.dtype))


def one_hot(labels: torch.Tensor, num_classes: int, device=None, dtype=None) -> torch.Tensor:
    if labels.dim() != 3 or labels.dtype != torch.int64:
        raise ValueError(f"labels must be an int64 tensor of shape BxHxW, got {labels.dtype} {tuple(labels.shape)}")
    B, H, W = labels.shape
    return torch.zeros(B, num_classes, H, W, device=device, dtype=dtype).scatter_(1, labels.unsqueeze(1), 1.0)


def _check(input: torch.Tensor, target: torch.Tensor) -> None:
    if input.dim() != 4 or input.shape[-2:] != target.shape[-2:] or input.device != target.device:
        raise ValueError(f"expected logits BxNxHxW and a target BxHxW on one device, got {tuple(input.shape)} / {tuple(target.shape)}")


def _dice(prob: torch.Tensor, hot: torch.Tensor, eps: float) -> torch.Tensor:
    """Mean over the batch of 1 - 2 <p, t> / (<p, p> + <t, 1> + eps) per class, then the mean of the two classes."""
    inter = (prob * hot).sum((2, 3))
    card = (prob * prob + hot).sum((2, 3))
    per_class = (1.0 - 2.0 * inter / (card + eps)).mean(0)
    return (per_class[1] + per_class[0]) / 2


class MultiClassDiceLoss(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.eps = 1e-6

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        _check(input, target)
        return _dice(F.softmax(input, dim=1), one_hot(target, input.shape[1], input.device, input.dtype), self.eps)


class DiceFocalLoss(nn.Module):
    def __init__(self, focal_rate=3, dice_rate=1) -> None:
        super().__init__()
        self.eps, self.focal_rate, self.dice_rate = 1e-6, focal_rate, dice_rate

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        _check(input, target)
        prob = F.softmax(input, dim=1)
        hot = one_hot(target, input.shape[1], input.device, input.dtype)
        pt = prob * hot + (1 - prob) * (1 - hot)
        focal = -(0.25 * (1 - pt).pow(2.0)) * (hot * torch.log(pt + 1e-5) + (1 - hot) * torch.log(1 - pt + 1e-5))
        return _dice(prob, hot, self.eps) * self.dice_rate + focal.mean() * self.focal_rate


class BoundaryLoss(nn.Module):
    def __init__(self, theta0=3, theta=5):
        super().__init__()
        self.theta0, self.theta = theta0, theta

    def forward(self, pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
        """pred: class probabilities (N, C, H, W); gt: one-hot target (N, C, H, W)."""
        n, c = pred.shape[:2]

        def edge(t):          # max-pool of the complement minus the complement = one-pixel inner boundary
            return F.max_pool2d(1 - t, self.theta0, stride=1, padding=(self.theta0 - 1) // 2) - (1 - t)

        def grow(t):
            return F.max_pool2d(t, self.theta, stride=1, padding=(self.theta - 1) // 2)
        gt_b, pred_b = edge(gt), edge(pred)
        gt_e, pred_e = grow(gt_b).view(n, c, -1), grow(pred_b).view(n, c, -1)
        gt_b, pred_b = gt_b.view(n, c, -1), pred_b.view(n, c, -1)
        P = (pred_b * gt_e).sum(2) / (pred_b.sum(2) + 1e-7)
        R = (pred_e * gt_b).sum(2) / (gt_b.sum(2) + 1e-7)
        return (1 - 2 * P * R / (P + R + 1e-7)).mean()


class DiceBoundaryLoss(nn.Module):
    def __init__(self, boundary_rate=0.05, dice_rate=1) -> None:
        super().__init__()
        self.eps, self.boundary_rate, self.dice_rate = 1e-6, boundary_rate, dice_rate
        self.BoundaryLoss = BoundaryLoss()

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        _check(input, target)
        prob = F.softmax(input, dim=1)
        hot = one_hot(target, input.shape[1], input.device, input.dtype)
        return _dice(prob, hot, self.eps) * self.dice_rate + self.BoundaryLoss(prob, hot) * self.boundary_rate


def build_criterion(args):
    """train.py:703-713: ``--loss mc_dice | dice_focal | dice_boundary`` or the weighted cross-entropy."""
    kind = getattr(args, "loss", "ce")
    if kind == "mc_dice":
        return MultiClassDiceLoss()
    if kind == "dice_focal":
        return DiceFocalLoss(getattr(args, "loss_focal_rate", 3), getattr(args, "loss_dice_rate", 1))
    if kind == "dice_boundary":
        return DiceBoundaryLoss(getattr(args, "loss_boundary_rate", 0.05), getattr(args, "loss_dice_rate", 1))
    return cross_entropy_loss
