"""Host-side mirror of the reference's ``utils/metrics.py``: same names and call conventions
(``f(logits[B,2,H,W], true[B,H,W] or [B,1,H,W]) -> 0-dim tensor`` with ``.backward()`` / ``.item()``), computed by the
fused CUDA loss kernels (one reduction pass + one elementwise gradient pass instead of ~10 ATen kernels).

The ``dims`` quirk of the reference is kept: with 3-D labels (what ``train.py:85,92`` passes) the soft sums run over
batch and H only and the ratio is averaged over (class, column); with 4-D labels over (class).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _SegLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, kind, alpha, beta, gamma, eps):
        loss, dlogits = ops.seg_loss_fwd_bwd(kind, logits.detach().contiguous(), labels.contiguous(), alpha, beta, gamma, eps)
        ctx.save_for_backward(dlogits)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return dlogits * g, None, None, None, None, None, None


def _loss(kind, logits, true, alpha=0.5, beta=0.5, gamma=0.0, eps=1e-7):
    if not logits.is_cuda:
        raise RuntimeError("fabric_b200 losses run on sm_100 CUDA tensors only (no CPU fallback)")
    if logits.shape[1] != 2:
        raise NotImplementedError("fabric_b200 losses are built for the reference's n_classes == 2")
    return _SegLoss.apply(logits.float(), true, kind, alpha, beta, gamma, eps)


def dice_loss(logits, true, eps=1e-7):
    """reference utils/metrics.py:51-83"""
    return _loss("dice", logits, true, eps=eps)


def jaccard_loss(logits, true, eps=1e-7):
    """reference utils/metrics.py:86-119"""
    return _loss("jaccard", logits, true, eps=eps)


class TverskyLoss(nn.Module):
    """reference utils/metrics.py:122-171 (default loss, metadata.json:42-44: alpha 0.1, beta 0.9)"""

    def __init__(self, alpha=0.5, beta=0.5, eps=1e-7, size_average=True):
        super(TverskyLoss, self).__init__()
        self.alpha = alpha
        self.beta = beta
        self.size_average = size_average
        self.eps = eps

    def forward(self, logits, true):
        return _loss("tversky", logits, true, self.alpha, self.beta, eps=self.eps)


class FocalLoss(nn.Module):
    """reference utils/metrics.py:8-48 with alpha=None, size_average=True (what helpers.py:306 builds)"""

    def __init__(self, gamma=0, alpha=None, size_average=True):
        super(FocalLoss, self).__init__()
        if alpha is not None or not size_average:
            raise NotImplementedError("the reference only builds FocalLoss(gamma) (utils/helpers.py:306)")
        self.gamma = gamma
        self.alpha = alpha
        self.size_average = size_average

    def forward(self, input, target):
        return _loss("focal", input, target, gamma=self.gamma)


def cross_entropy_loss(logits, true):
    """2-class softmax cross entropy: the working stand-in for the reference's `bce` choice, which raises on
    2-channel logits (utils/helpers.py:303-304)."""
    return _loss("ce", logits, true)


def get_criterion(opt):
    """reference utils/helpers.py:288-314 (same `opt` attributes)"""
    if opt.loss_function == 'bce':
        return cross_entropy_loss
    if opt.loss_function == 'focal':
        return FocalLoss(opt.focal_gamma)
    if opt.loss_function == 'dice':
        return dice_loss
    if opt.loss_function == 'jaccard':
        return jaccard_loss
    if opt.loss_function == 'tversky':
        return TverskyLoss(alpha=opt.tversky_alpha, beta=opt.tversky_beta)
    raise ValueError(opt.loss_function)
