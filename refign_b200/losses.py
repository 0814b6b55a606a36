"""Flow losses of the alignment network's own training: the reference's
``MultiScaleFlowLoss`` / ``WBipathLoss`` / ``HuberLoss`` interfaces (reference models/losses.py:25-328,
constructor arguments and ``forward`` signatures kept so the YAML ``class_path`` entries switch over), restated on
top of this package's ``warp`` (CUDA kernel with backward).

Both losses work on the head's coarse-to-fine list of ``(flow, log-variance)`` levels.  All reductions are
masked means; an empty mask gives a zero loss for that level (losses.py:96-98)."""
import math
from collections.abc import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .matching_utils import warp


def _resize(x, size):
    return F.interpolate(x, size, mode='bilinear', align_corners=False)


def _mask_at(mask, size):
    """[B,H,W] validity mask -> bool [B,1,h,w]; a pixel stays valid only if every contributing pixel is
    (bilinear resize of the 0/1 mask followed by floor, losses.py:92-95)."""
    mask = mask.unsqueeze(1)
    if tuple(mask.shape[-2:]) != tuple(size):
        mask = _resize(mask.float(), size).floor().bool()
    return mask.bool()


def flow_in_image_mask(flow):
    """True where ``pixel + flow`` lands inside the image (reference helpers/matching_utils.py:60-74)."""
    B, _, H, W = flow.shape
    xs = torch.arange(W, dtype=flow.dtype, device=flow.device).view(1, 1, W)
    ys = torch.arange(H, dtype=flow.dtype, device=flow.device).view(1, H, 1)
    mx, my = flow[:, 0] + xs, flow[:, 1] + ys
    return (mx >= 0) & (mx <= W - 1) & (my >= 0) & (my <= H - 1)


class HuberLoss(nn.Module):
    """2 * delta * smooth_l1 (the factor makes it the negative log-likelihood of the probabilistic set-up,
    losses.py:25-35)."""

    def __init__(self, reduction='mean', delta=1.0):
        super().__init__()
        self.reduction, self.delta = reduction, delta

    def forward(self, input, target):
        return 2.0 * self.delta * F.smooth_l1_loss(input, target, reduction=self.reduction, beta=self.delta)


class MultiScaleFlowLoss(nn.Module):
    """Weighted sum over pyramid levels of the masked-mean flow error; with a log-variance per level the
    error becomes the Laplace/Gaussian NLL ``0.5 exp(-s) e + s + log 2 pi`` (two variances of a composed flow
    are merged with logsumexp) -- losses.py:38-191."""

    _ERRORS = {'L1Loss': lambda: nn.L1Loss(reduction='none'), 'L2Loss': lambda: nn.MSELoss(reduction='none'),
               'HuberLoss': lambda: HuberLoss(reduction='none')}

    def __init__(self, level_weights=None, loss_type='L1Loss', downsample_gt_flow=True, reduction='mean'):
        super().__init__()
        if loss_type not in self._ERRORS:
            raise ValueError(loss_type)
        if reduction != 'mean':
            raise ValueError(reduction)
        self.level_weights = level_weights
        self.downsample_gt_flow = downsample_gt_flow
        self.reduction = reduction
        self.loss_type = loss_type
        self.loss_function = self._ERRORS[loss_type]()

    def _level(self, flow, logvar, gt_flow, mask):
        if self.downsample_gt_flow:
            size = flow.shape[-2:]
            gt_flow = _resize(gt_flow, size)
        else:
            size = gt_flow.shape[-2:]
            flow = _resize(flow, size)
            if logvar is not None:
                logvar = _resize(logvar, size)
        if mask is not None:
            mask = _mask_at(mask, size)
            if not bool(mask.any()):
                return flow.new_zeros([])
        err = self.loss_function(flow, gt_flow).sum(1, keepdim=True)
        if logvar is not None:
            if self.loss_type not in ('L2Loss', 'HuberLoss'):
                raise AssertionError("the probabilistic loss needs L2Loss or HuberLoss")
            if logvar.shape[1] == 2:
                logvar = torch.logsumexp(logvar, 1, keepdim=True)
            elif logvar.shape[1] != 1:
                raise ValueError("1 or 2 log-variance channels expected")
            err = 0.5 * torch.exp(-logvar) * err + logvar + math.log(2 * math.pi)
        return torch.masked_select(err, mask).mean()

    # reference method names (losses.py:71,125)
    def probabilistic_one_scale(self, est_flow, est_uncert, gt_flow, mask=None):
        return self._level(est_flow, est_uncert, gt_flow, mask)

    def one_scale(self, est_flow, gt_flow, mask=None):
        return self._level(est_flow, None, gt_flow, mask)

    def forward(self, flow_output, gt_flow, mask=None):
        levels = list(flow_output) if isinstance(flow_output, Sequence) else [flow_output]
        weights = self.level_weights if self.level_weights else [1] * len(levels)
        assert len(weights) == len(levels)
        total = 0
        for i, (level, wgt) in enumerate(zip(levels, weights)):
            m = mask[i] if (mask is not None and isinstance(mask, Sequence)) else mask
            flow, logvar = level if isinstance(level, tuple) else (level, None)
            total = total + wgt * self._level(flow, logvar, gt_flow, m)
        return total


class WBipathLoss(nn.Module):
    """W-bipath constraint: the flow prime->source composed with source->target (the latter warped by the
    former) must equal the synthetic flow prime->target (losses.py:194-328)."""

    def __init__(self, objective='multi_scale_flow_loss', reduction='mean', level_weights=None, loss_type='L1Loss',
                 downsample_gt_flow=True, detach_flow_for_warping=True, visibility_mask=False, alpha_1=0.03,
                 alpha_2=0.5):
        super().__init__()
        if objective != 'multi_scale_flow_loss':
            raise ValueError(objective)
        self.objective = MultiScaleFlowLoss(level_weights=level_weights, loss_type=loss_type,
                                            downsample_gt_flow=downsample_gt_flow, reduction=reduction)
        self.detach_flow_for_warping = detach_flow_for_warping
        self.visibility_mask = visibility_mask
        self.alpha_1, self.alpha_2 = alpha_1, alpha_2

    @staticmethod
    def length_sq(x):
        return (x ** 2).sum(1)

    @torch.no_grad()
    def get_cyclic_consistency_mask(self, flow_prime_to_source, warped_flow_source_to_target, synthetic_flow):
        """Forward-backward visibility test (losses.py:236-253): not occluded where the composition error is
        below ``alpha_1 * (|a|^2 + |b|^2 + |gt|^2) + alpha_2``."""
        gt = _resize(synthetic_flow, flow_prime_to_source.shape[-2:])
        bound = self.alpha_1 * (self.length_sq(flow_prime_to_source) + self.length_sq(warped_flow_source_to_target)
                                + self.length_sq(gt)) + self.alpha_2
        return ~(self.length_sq(flow_prime_to_source + warped_flow_source_to_target - gt) > bound)

    def forward(self, estimated_flow_target_prime_to_source, estimated_flow_source_to_target, flow_map, mask_used,
                return_masks=False):
        H, W = flow_map.shape[-2:]
        a_levels = estimated_flow_target_prime_to_source
        b_levels = estimated_flow_source_to_target
        if not isinstance(a_levels, Sequence):
            a_levels = [a_levels]
        if not isinstance(b_levels, Sequence):
            b_levels = [b_levels]
        composed, masks, cyc = [], [], []
        for a, b in zip(a_levels, b_levels):
            (fa, ua), (fb, ub) = (a, b) if isinstance(a, tuple) else ((a, None), (b, None))
            h, w = fa.shape[-2:]
            # the warping flow is expressed in pixels of THIS level (the estimates are in image pixels)
            wf = fa.detach() if self.detach_flow_for_warping else fa
            wf = torch.stack((wf[:, 0] * (float(w) / float(W)), wf[:, 1] * (float(h) / float(H))), 1)
            fb_w = warp(fb, wf)
            flow = fa + fb_w
            composed.append((flow, torch.cat((ua, warp(ub, wf)), 1)) if ua is not None else flow)
            m = flow_in_image_mask(wf.detach())
            if mask_used is not None:
                m = m & _resize(mask_used.unsqueeze(1).float(), (h, w)).squeeze(1).floor().bool()
            if self.visibility_mask:
                mc = self.get_cyclic_consistency_mask(fa.detach(), fb_w.detach(), flow_map)
                m = m & mc
                cyc.append(mc)
            masks.append(m)
        loss = self.objective(composed, flow_map, mask=masks)
        if return_masks:
            return loss, masks, (cyc if cyc else None), composed
        return loss
