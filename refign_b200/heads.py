"""Decode heads with the reference's interface and ``state_dict`` keys:
``DAFormerHead`` (reference models/heads/daformer.py) and ``UAWarpCHead``
(reference models/heads/uawarpc.py), plus ``BaseHead`` (models/heads/base.py).

The UAWarpC head is where the hand-written kernels sit: global correlation (+ mutual matching,
ReLU, L2-norm), three levels of 9x9 local correlation with the ReLU/L2-norm epilogue fused, and the
bilinear feature warps -- all without the host synchronisation the reference's ``warp`` performs
(helpers/matching_utils.py:19).
"""
import math
import os
from collections.abc import Iterable

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .matching_utils import unnormalise_and_convert_mapping_to_flow, warp
from .modules import (MLP, ConvBNReLU, GlobalFeatureCorrelationLayer, LocalFeatureCorrelationLayer,
                      OpticalFlowEstimatorResidualConnection, RefinementModule, UncertaintyModule)


class BaseHead(nn.Module):
    def __init__(self, num_classes, in_index, input_transform=None):
        super().__init__()
        self.input_transform = input_transform
        self.in_index = in_index[0] if isinstance(in_index, Iterable) and len(in_index) == 1 else in_index
        self.num_classes = num_classes

    def forward(self, inp):
        raise NotImplementedError

    def _transform_inputs(self, inputs):
        if self.input_transform == 'resize_concat':
            sel = [inputs[i] for i in self.in_index]
            return torch.cat([F.interpolate(x, size=sel[0].shape[2:], mode='bilinear', align_corners=False)
                              for x in sel], dim=1)
        if self.input_transform == 'multiple_select':
            return [inputs[i] for i in self.in_index]
        return inputs[self.in_index]


# ------------------------------------------------------------------------------------------------
# DAFormer
# ------------------------------------------------------------------------------------------------
class _ASPP(nn.Module):
    """Depthwise-separable ASPP + 3x3 bottleneck; attribute names follow the reference's
    ASPPWrapper (daformer.py:65-126) so checkpoints load: ``aspp_modules.{i}``, ``bottleneck``."""

    def __init__(self, in_channels, channels, dilations, sep=True, norm_layer=nn.BatchNorm2d,
                 activation_layer=nn.ReLU):
        super().__init__()
        self.dilations = tuple(dilations)
        branches = []
        for d in self.dilations:
            if d == 1:
                branches.append(ConvBNReLU(in_channels, channels, 1, padding=0, norm_layer=norm_layer,
                                           activation_layer=activation_layer))
            else:
                branches.append(ConvBNReLU(in_channels, channels, 3, dilation=d, padding=d, norm_layer=norm_layer,
                                           activation_layer=activation_layer, depthwise_separable=sep))
        self.aspp_modules = nn.ModuleList(branches)
        self.bottleneck = ConvBNReLU(len(self.dilations) * channels, channels, 3, padding=1, norm_layer=norm_layer,
                                     activation_layer=activation_layer)

    def forward(self, x):
        return self.bottleneck(torch.cat([m(x) for m in self.aspp_modules], dim=1))


class DAFormerHead(BaseHead):
    def __init__(self, in_channels, in_index, num_classes, input_transform=None, channels=256, dropout_ratio=0.1,
                 embed_dims=256):
        super().__init__(num_classes, in_index, input_transform)
        self.in_channels = in_channels
        self.channels = channels
        if isinstance(embed_dims, int):
            embed_dims = [embed_dims] * len(in_channels)
        self.embed_layers = nn.ModuleDict({str(i): MLP(input_dim=c, embed_dim=e)
                                           for i, (c, e) in enumerate(zip(in_channels, embed_dims))})
        self.fuse_layer = _ASPP(sum(embed_dims), channels, dilations=(1, 6, 12, 18), sep=True)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        self.conv_seg = nn.Conv2d(channels, num_classes, kernel_size=1)
        nn.init.normal_(self.conv_seg.weight, mean=0, std=0.01)
        nn.init.zeros_(self.conv_seg.bias)
        for m in self.modules():  # mmseg-style init of the plain conv blocks (daformer.py:188-201)
            if isinstance(m, ConvBNReLU) and not m.depthwise_separable:
                nn.init.kaiming_normal_(m.conv.weight, a=0, mode='fan_out', nonlinearity='relu')
                if m.conv.bias is not None:
                    nn.init.zeros_(m.conv.bias)
                if m.use_norm:
                    nn.init.ones_(m.bn.weight)
                    nn.init.zeros_(m.bn.bias)

    def forward(self, x):
        x = self._transform_inputs(x)
        n = x[-1].shape[0]
        os_size = x[0].shape[2:]
        cl = x[0].is_cuda
        embedded, sizes = [], []
        for i in range(len(self.in_channels)):
            hi, wi = x[i].shape[2:]
            embedded.append(self.embed_layers[str(i)](x[i]))    # [n, hi*wi, E] tokens
            sizes.append((hi, wi))
        if (cl and len(embedded) <= 4 and all(c.dtype == torch.bfloat16 and c.shape[-1] % 8 == 0 for c in embedded)
                and all(hi <= os_size[0] and wi <= os_size[1] for hi, wi in sizes)):
            # one kernel writes the concatenated channels-last tensor (resize + cat of reference :203-221)
            y = ops.upsample_concat(embedded, sizes, os_size)
        else:
            maps = []
            for c, (hi, wi) in zip(embedded, sizes):
                c = c.view(n, hi, wi, -1).permute(0, 3, 1, 2)       # NCHW view, channels-last memory
                if (hi, wi) != tuple(os_size):
                    # autocast would promote the upsampling to fp32 (a 4-byte [n,256,H/4,W/4] tensor per stage and
                    # an fp32 ASPP input); the kernel accumulates in fp32 either way, keep the activations bf16
                    with torch.autocast('cuda', enabled=False):
                        c = F.interpolate(c, size=os_size, mode='bilinear', align_corners=False)
                maps.append(c)
            y = torch.cat(maps, dim=1)
            if cl:
                y = y.contiguous(memory_format=torch.channels_last)
        y = self.fuse_layer(y)
        if self.dropout is not None:
            y = self.dropout(y)
        return self.conv_seg(y)


# ------------------------------------------------------------------------------------------------
# SegFormer (the HRDA scale-attention head: configs/*/refign_hrda_star.yaml `hrda_scale_attention`)
# ------------------------------------------------------------------------------------------------
def _embed_resize_concat(x, embed, order, os_size):
    """``cat([resize(embed_i(x_i)) for i in order], 1)``: the per-stage linear embedding (tokens), bilinear
    resize to the stride-4 map and channel concat shared by the DAFormer and SegFormer heads.  On the GPU with
    bf16 tokens this is ONE kernel writing the channels-last concatenation (ops.upsample_concat)."""
    n = x[0].shape[0]
    embedded, sizes = [], []
    for i in order:
        hi, wi = x[i].shape[2:]
        embedded.append(embed(i)(x[i]))    # [n, hi*wi, E] tokens
        sizes.append((hi, wi))
    cl = x[0].is_cuda
    if (cl and len(embedded) <= 4 and all(c.dtype == torch.bfloat16 and c.shape[-1] % 8 == 0 for c in embedded)
            and all(hi <= os_size[0] and wi <= os_size[1] for hi, wi in sizes)):
        return ops.upsample_concat(embedded, sizes, os_size)
    maps = []
    for c, (hi, wi) in zip(embedded, sizes):
        c = c.view(n, hi, wi, -1).permute(0, 3, 1, 2)       # NCHW view, channels-last memory
        if (hi, wi) != tuple(os_size):
            with torch.autocast('cuda', enabled=False):
                c = F.interpolate(c, size=os_size, mode='bilinear', align_corners=False)
        maps.append(c)
    y = torch.cat(maps, dim=1)
    return y.contiguous(memory_format=torch.channels_last) if cl else y


class SegFormerHead(BaseHead):
    """All-MLP decoder (reference models/heads/segformer.py:15-111; same ``state_dict`` keys:
    ``linear_c{1..4}.proj``, ``linear_fuse.{conv,bn}``, ``linear_pred``): per-stage linear embedding, resize to
    stride 4, concat in the order c4, c3, c2, c1, 1x1 conv + BN + ReLU, 1x1 classifier."""

    os = 4

    def __init__(self, in_channels, in_index, num_classes, input_transform=None, channels=256, dropout_ratio=0.1):
        super().__init__(num_classes, in_index, input_transform)
        self.in_channels = in_channels
        c1, c2, c3, c4 = in_channels
        self.linear_c4 = MLP(input_dim=c4, embed_dim=channels)
        self.linear_c3 = MLP(input_dim=c3, embed_dim=channels)
        self.linear_c2 = MLP(input_dim=c2, embed_dim=channels)
        self.linear_c1 = MLP(input_dim=c1, embed_dim=channels)
        self.linear_fuse = ConvBNReLU(channels * 4, channels, 1, norm_layer=nn.BatchNorm2d)
        self.linear_pred = nn.Conv2d(channels, num_classes, kernel_size=1)
        self.dropout = nn.Dropout2d(dropout_ratio) if dropout_ratio > 0 else None
        nn.init.normal_(self.linear_pred.weight, mean=0, std=0.01)
        nn.init.zeros_(self.linear_pred.bias)
        nn.init.kaiming_normal_(self.linear_fuse.conv.weight, a=0, mode='fan_out', nonlinearity='relu')
        nn.init.ones_(self.linear_fuse.bn.weight)
        nn.init.zeros_(self.linear_fuse.bn.bias)

    def forward(self, inputs):
        c = list(inputs)   # the reference unpacks the four stage maps directly (segformer.py:79)
        embeds = (self.linear_c1, self.linear_c2, self.linear_c3, self.linear_c4)
        y = _embed_resize_concat(c, lambda i: embeds[i], (3, 2, 1, 0), c[0].shape[2:])
        y = self.linear_fuse(y)
        if self.dropout is not None:
            y = self.dropout(y)
        return self.linear_pred(y)


# ------------------------------------------------------------------------------------------------
# UAWarpC
# ------------------------------------------------------------------------------------------------
def _l2n(t):
    return F.normalize(t.float(), p=2, dim=1)


def _up(x, size):
    return F.interpolate(x, size=size, mode='bilinear', align_corners=False)


def _scale_flow(flow, sx, sy):
    """flow * (sx, sy) per channel without the reference's in-place clone dance (and without a
    host->device tensor creation, which a CUDA-graph capture of the step would not allow)."""
    if sx == sy:
        return flow * sx
    return torch.cat((flow[:, 0:1] * sx, flow[:, 1:2] * sy), dim=1)


class UAWarpCHead(BaseHead):
    """4-level coarse-to-fine flow + log-variance (reference uawarpc.py:17-305)."""

    def __init__(self, in_index, input_transform=None, pretrained=None, batch_norm=True,
                 refinement_at_adaptive_res=True, refinement_at_finest_level=True, estimate_uncertainty=False,
                 uncertainty_mixture=False, iterative_refinement=False):
        super().__init__(None, in_index, input_transform)
        self.estimate_uncertainty = estimate_uncertainty
        self.uncertainty_mixture = uncertainty_mixture
        self.iterative_refinement = iterative_refinement
        self.global_corr = GlobalFeatureCorrelationLayer(cyclic_consistency=True)
        self.local_corr = LocalFeatureCorrelationLayer(patch_size=9)
        u = 1 if estimate_uncertainty else 0
        dec = lambda c: OpticalFlowEstimatorResidualConnection(in_channels=c, batch_norm=batch_norm, output_x=True)
        self.decoder4 = dec(16 * 16)
        self.decoder3 = dec(81 + 2 + u)
        self.refinement_at_adaptive_res = refinement_at_adaptive_res
        if refinement_at_adaptive_res:
            self.refinement_module_adaptive = RefinementModule(32, batch_norm=batch_norm)
        self.decoder2 = dec(81 + 2 + u)
        self.reduce = nn.Conv2d(32, 2, kernel_size=1, bias=True)
        self.decoder1 = dec(81 + 2 + 2 + u)
        self.refinement_at_finest_level = refinement_at_finest_level
        if refinement_at_finest_level:
            self.refinement_module_finest = RefinementModule(32, batch_norm=batch_norm)
        if estimate_uncertainty:
            self.estimate_uncertainty_components4 = UncertaintyModule(in_channels=1, search_size=16)
            for lvl in (3, 2, 1):
                setattr(self, 'estimate_uncertainty_components%d' % lvl,
                        UncertaintyModule(in_channels=1, search_size=9, feed_in_previous=True))
        if pretrained is not None:
            self.load_weights(pretrained)

    def load_weights(self, pretrain_path):
        if pretrain_path is None:
            return
        from .mix_transformer import resolve_checkpoint
        ckpt = torch.load(resolve_checkpoint(pretrain_path), map_location='cpu')
        sd = ckpt['state_dict'] if 'state_dict' in ckpt else ckpt
        pre = 'alignment_head.'
        self.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=True)

    def _level(self, lvl, c_src, c_trg, up_flow, up_u, res_scale, extra=None):
        """One local-correlation level: warp source by the (rescaled) flow, correlate, decode."""
        h, w = c_trg.shape[-2:]
        warped = warp(c_src, _scale_flow(up_flow, *res_scale))
        corr = self.local_corr(warped, c_trg)
        parts = [corr, up_flow]
        if extra is not None:
            parts.append(extra)
        if self.estimate_uncertainty:
            parts.append(up_u)
        res, feat = getattr(self, 'decoder%d' % lvl)(torch.cat(parts, 1))
        return corr, res, feat

    def forward(self, trg, src, trg_256, src_256, out_size, **kwargs):
        c11, c12 = (_l2n(t) for t in self._transform_inputs(trg))
        c13, c14 = (_l2n(t) for t in self._transform_inputs(trg_256))
        c21, c22 = (_l2n(t) for t in self._transform_inputs(src))
        c23, c24 = (_l2n(t) for t in self._transform_inputs(src_256))
        h0, w0 = out_size
        unc = self.estimate_uncertainty
        # ---- level 4: 16x16 global correlation, flow in 256-px units -----------------------------
        assert c14.shape[-2:] == (16, 16), c14.shape
        corr4 = self.global_corr(c24, c14)
        map4, x4 = self.decoder4(corr4)
        flow4 = unnormalise_and_convert_mapping_to_flow(map4) * (256.0 / 16.0)
        u4 = None
        if unc:
            u4 = self.estimate_uncertainty_components4(corr4, x4) + 2.0 * math.log(256.0 / 16.0)
        # ---- level 3: 32x32 ---------------------------------------------------------------------
        assert c13.shape[-2:] == (32, 32), c13.shape
        up_flow4 = _up(flow4, (32, 32))
        up_u4 = _up(u4, (32, 32)) if unc else None
        corr3, res3, x3 = self._level(3, c23, c13, up_flow4, up_u4, (32.0 / 256.0, 32.0 / 256.0))
        if self.refinement_at_adaptive_res:
            res3 = res3 + self.refinement_module_adaptive(x3)
        flow3 = res3 + up_flow4
        u3 = self.estimate_uncertainty_components3(corr3, x3, up_u4, up_flow4) if unc else None
        # to original-resolution pixel units (uawarpc.py:163-173)
        flow3 = _scale_flow(flow3, float(w0) / 256.0, float(h0) / 256.0)
        diag = math.sqrt(h0 ** 2 + w0 ** 2) / math.sqrt(256.0 ** 2 + 256.0 ** 2)
        if unc:
            u3 = u3 + 2.0 * math.log(diag)
        if self.iterative_refinement and not self.training:
            raise NotImplementedError("iterative_refinement (eval-only, off in every Refign UDA config) "
                                      "is not implemented")
        # ---- level 2: 1/8 -----------------------------------------------------------------------
        h2, w2 = c12.shape[-2:]
        up_flow3 = _up(flow3, (h2, w2))
        up_u3 = _up(u3, (h2, w2)) if unc else None
        corr2, res2, x2 = self._level(2, c22, c12, up_flow3, up_u3, (w2 / float(w0), h2 / float(h0)))
        flow2 = res2 + up_flow3
        u2 = self.estimate_uncertainty_components2(corr2, x2, up_u3, up_flow3) if unc else None
        # ---- level 1: 1/4 -----------------------------------------------------------------------
        h1, w1 = c11.shape[-2:]
        up_flow2 = _up(flow2, (h1, w1))
        up_u2 = _up(u2, (h1, w1)) if unc else None
        up_feat2 = self.reduce(_up(x2, (h1, w1)))
        corr1, res1, x1 = self._level(1, c21, c11, up_flow2, up_u2, (w1 / float(w0), h1 / float(h0)), up_feat2)
        if self.refinement_at_finest_level:
            res1 = res1 + self.refinement_module_finest(x1)
        flow1 = res1 + up_flow2
        u1 = self.estimate_uncertainty_components1(corr1, x1, up_u2, up_flow2) if unc else None
        # level-4 outputs are returned in original-resolution units too (uawarpc.py:264-276)
        flow4 = _scale_flow(flow4, float(w0) / 256.0, float(h0) / 256.0)
        if unc:
            u4 = u4 + 2.0 * math.log(diag)
            return (flow4, u4), (flow3, u3), (flow2, u2), (flow1, u1)
        return flow4, flow3, flow2, flow1
