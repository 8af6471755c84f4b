"""Detection losses used inside the training step (reference layers/losses/det3d.py:10-34,279-322,364-421).

Same math; the normaliser all-reduces (`reduce_mean`, one per call in the reference -> 24 per step) can
be supplied pre-reduced through the optional `num_pos` / `num` arguments so a step needs a single fused
all-reduce of all GT-only normalisers.  Without them the reference behaviour is kept.
"""
import torch
import torch.nn as nn

from unidistill.utils.torch_dist import reduce_mean


class AutomaticWeightedLoss(nn.Module):
    """Uncertainty-weighted sum: sum_i 0.5 / p_i^2 * loss_i + log(1 + p_i^2)."""

    def __init__(self, num=2):
        super().__init__()
        self.params = nn.Parameter(torch.ones(num, requires_grad=True))

    def forward(self, *losses):
        total = 0
        for i, loss in enumerate(losses):
            p2 = self.params[i] ** 2
            total = total + 0.5 / p2 * loss + torch.log(1 + p2)
        return total


class FocalLoss(nn.Module):
    """CornerNet-style penalty-reduced focal loss on a clamped-sigmoid heat-map."""

    def __init__(self, alpha, gamma):
        super().__init__()
        self.alpha = alpha
        self.gamma = gamma

    def _neg_loss(self, pred, gt, num_pos=None):
        pos = gt.eq(1).to(pred.dtype)
        neg = gt.eq(0).to(pred.dtype)
        pos_loss = (torch.log(pred) * torch.pow(1 - pred, self.gamma) * pos * self.alpha).sum()
        neg_loss = (torch.log(1 - pred + 1e-4) * torch.pow(pred, self.gamma) * neg * (1 - self.alpha)).sum()
        if num_pos is None:
            num_pos = reduce_mean(pos.sum())
        # sync-free form of:  -neg_loss if num_pos == 0 else -(pos_loss + neg_loss) / num_pos
        # (with num_pos == 0 there are no positives, so pos_loss == 0 as well)
        safe = torch.where(num_pos == 0, torch.ones_like(num_pos), num_pos)
        return -(pos_loss + neg_loss) / safe

    def channel_sums(self, pred, gt):
        """(B,C,H,W) -> (C,): pos_loss + neg_loss of `_neg_loss` summed per channel (the batched head loss adds the
        channels of each task and divides by its own normaliser)."""
        pos = gt.eq(1).to(pred.dtype)
        neg = gt.eq(0).to(pred.dtype)
        both = torch.log(pred) * torch.pow(1 - pred, self.gamma) * pos * self.alpha \
            + torch.log(1 - pred + 1e-4) * torch.pow(pred, self.gamma) * neg * (1 - self.alpha)
        return both.sum(dim=(0, 2, 3))

    def forward(self, out, target, num_pos=None):
        return self._neg_loss(out, target, num_pos)


class CenterNetFocalLoss(nn.Module):
    """The head's default `crit` (reference layers/losses/det3d.py:325-360): CornerNet focal loss for Gaussian
    (non-binary) heat-maps -- no alpha weighting, every cell with gt < 1 is a negative weighted by (1 - gt)^4, and
    log(1 - pred) carries no epsilon.  Not the same function as `FocalLoss` above."""

    def _neg_loss(self, pred, gt, num_pos=None):
        pos = gt.eq(1).to(pred.dtype)
        neg = gt.lt(1).to(pred.dtype)
        pos_loss = (torch.log(pred) * torch.pow(1 - pred, 2) * pos).sum()
        neg_loss = (torch.log(1 - pred) * torch.pow(pred, 2) * torch.pow(1 - gt, 4) * neg).sum()
        if num_pos is None:
            num_pos = reduce_mean(pos.sum())
        # sync-free form of the reference's `if num_pos == 0` (no positives => pos_loss == 0)
        safe = torch.where(num_pos == 0, torch.ones_like(num_pos), num_pos)
        return -(pos_loss + neg_loss) / safe

    def channel_sums(self, pred, gt):
        """(B,C,H,W) -> (C,): pos_loss + neg_loss of `_neg_loss` summed per channel."""
        pos = gt.eq(1).to(pred.dtype)
        neg = gt.lt(1).to(pred.dtype)
        both = torch.log(pred) * torch.pow(1 - pred, 2) * pos \
            + torch.log(1 - pred) * torch.pow(pred, 2) * torch.pow(1 - gt, 4) * neg
        return both.sum(dim=(0, 2, 3))

    def forward(self, out, target, num_pos=None):
        return self._neg_loss(out, target, num_pos)


def _gather_feat(feat, ind, mask=None):
    dim = feat.size(2)
    ind = ind.unsqueeze(2).expand(ind.size(0), ind.size(1), dim)
    feat = feat.gather(1, ind)
    if mask is not None:
        mask = mask.unsqueeze(2).expand_as(feat)
        feat = feat[mask].view(-1, dim)
    return feat


def _transpose_and_gather_feat(feat, ind):
    """(B,C,H,W), ind (B,K) flat positions -> (B,K,C)."""
    b, c = feat.size(0), feat.size(1)
    flat = feat.view(b, c, -1)
    return flat.gather(2, ind.unsqueeze(1).expand(b, c, ind.size(1))).permute(0, 2, 1).contiguous()


class CenterNetRegLoss(nn.Module):
    """Masked L1 regression at the gathered positions; returns one value per code dimension."""

    def _reg_loss(self, regr, gt_regr, mask, num=None):
        if num is None:
            num = reduce_mean(mask.float().sum())
        m = mask.unsqueeze(2).expand_as(gt_regr).float() * (~torch.isnan(gt_regr)).float()
        loss = torch.abs(regr * m - gt_regr * m).sum(dim=(0, 1))
        return loss / (num + 1e-4)

    def forward(self, output, mask, ind, target, num=None):
        return self._reg_loss(_transpose_and_gather_feat(output, ind), target, mask, num)
