"""Functional torch-CPU restatement of the network pieces of the Refign hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function takes a
reference-compatible ``state_dict`` (same keys/shapes as the reference modules, so
weights can be shared with the real reference and with refign_b200's modules) and
restates the cited reference forward in plain fp32 torch ops; the correlation, warp,
confidence and refinement operators come from the C oracle (oracle/refign_oracle.c).
Pinned against the real reference by tests/test_oracle_model_vs_reference.py and the
fixtures in tests/golden/model_*.npz.

Stochastic pieces (DropPath, Dropout2d, DACS colour jitter / blur, class-mix choice)
are parameters here: the parity configuration runs with their rates at zero
(SURVEY.md section 8d).
"""
import math

import torch
import torch.nn.functional as F

from . import cert as o_cert
from . import global_corr as o_global_corr
from . import local_corr_layer as o_local_corr_layer
from . import refine as o_refine
from . import warp as o_warp

MIT_SETTINGS = {  # models/backbones/mix_transformer.py:247-308
    "mit_b0": dict(embed_dims=[32, 64, 160, 256], depths=[2, 2, 2, 2]),
    "mit_b1": dict(embed_dims=[64, 128, 320, 512], depths=[2, 2, 2, 2]),
    "mit_b2": dict(embed_dims=[64, 128, 320, 512], depths=[3, 4, 6, 3]),
    "mit_b3": dict(embed_dims=[64, 128, 320, 512], depths=[3, 4, 18, 3]),
    "mit_b4": dict(embed_dims=[64, 128, 320, 512], depths=[3, 8, 27, 3]),
    "mit_b5": dict(embed_dims=[64, 128, 320, 512], depths=[3, 6, 40, 3]),
}
MIT_HEADS = [1, 2, 5, 8]
MIT_SR = [8, 4, 2, 1]


class SD:
    """state_dict view with a key prefix."""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def __call__(self, name):
        t = self.sd[self.prefix + name]      # not detached: oracle.train_step differentiates through these
        return t if t.dtype == torch.float32 else t.float()

    def has(self, name):
        return (self.prefix + name) in self.sd

    def sub(self, p):
        return SD(self.sd, self.prefix + p)


# ----------------------------------------------------------------------------
# MiT encoder  (models/backbones/mix_transformer.py)
# ----------------------------------------------------------------------------
def patch_embed(p, x, k, s):
    """OverlapPatchEmbed.forward (:236-242): conv k/s/pad k//2 -> NLC -> LayerNorm(1e-5)."""
    x = F.conv2d(x, p("proj.weight"), p("proj.bias"), stride=s, padding=k // 2)
    B, C, H, W = x.shape
    x = x.flatten(2).transpose(1, 2)
    return F.layer_norm(x, (C,), p("norm.weight"), p("norm.bias"), 1e-5), H, W


def sr_attention_core(q, k, v, scale):
    """softmax(q k^T scale) v per head (:156-160). q [B,h,N,d], k,v [B,h,M,d]."""
    attn = (q @ k.transpose(-2, -1)) * scale
    return attn.softmax(dim=-1) @ v


def attention(p, x, H, W, heads, sr):
    """Attention.forward (:137-164)."""
    B, N, C = x.shape
    d = C // heads
    q = F.linear(x, p("q.weight"), p("q.bias")).reshape(B, N, heads, d).permute(0, 2, 1, 3)
    if sr > 1:
        x_ = x.permute(0, 2, 1).reshape(B, C, H, W)
        x_ = F.conv2d(x_, p("sr.weight"), p("sr.bias"), stride=sr).reshape(B, C, -1).permute(0, 2, 1)
        x_ = F.layer_norm(x_, (C,), p("norm.weight"), p("norm.bias"), 1e-5)
    else:
        x_ = x
    kv = F.linear(x_, p("kv.weight"), p("kv.bias")).reshape(B, -1, 2, heads, d).permute(2, 0, 3, 1, 4)
    o = sr_attention_core(q, kv[0], kv[1], d ** -0.5)
    o = o.transpose(1, 2).reshape(B, N, C)
    return F.linear(o, p("proj.weight"), p("proj.bias"))


def mix_ffn(p, x, H, W):
    """Mlp.forward + DWConv (:96-103, :562-568); exact-erf GELU."""
    B, N, _ = x.shape
    x = F.linear(x, p("fc1.weight"), p("fc1.bias"))
    C = x.shape[-1]
    x = x.transpose(1, 2).reshape(B, C, H, W)
    x = F.conv2d(x, p("dwconv.dwconv.weight"), p("dwconv.dwconv.bias"), padding=1, groups=C)
    x = F.gelu(x.flatten(2).transpose(1, 2))
    return F.linear(x, p("fc2.weight"), p("fc2.bias"))


def mit_forward(sd, x, model_type="mit_b5", prefix=""):
    """MixVisionTransformer.forward_features (:511-547), drop_path = 0. Returns 4 NCHW maps."""
    p = SD(sd, prefix)
    cfg = MIT_SETTINGS[model_type]
    outs = []
    B = x.shape[0]
    for st in range(4):
        dim = cfg["embed_dims"][st]
        x, H, W = patch_embed(p.sub("patch_embed%d." % (st + 1)), x, 7 if st == 0 else 3, 4 if st == 0 else 2)
        for j in range(cfg["depths"][st]):
            b = p.sub("block%d.%d." % (st + 1, j))
            y = F.layer_norm(x, (dim,), b("norm1.weight"), b("norm1.bias"), 1e-6)
            x = x + attention(b.sub("attn."), y, H, W, MIT_HEADS[st], MIT_SR[st])
            y = F.layer_norm(x, (dim,), b("norm2.weight"), b("norm2.bias"), 1e-6)
            x = x + mix_ffn(b.sub("mlp."), y, H, W)
        x = F.layer_norm(x, (dim,), p("norm%d.weight" % (st + 1)), p("norm%d.bias" % (st + 1)), 1e-6)
        x = x.reshape(B, H, W, dim).permute(0, 3, 1, 2).contiguous()
        outs.append(x)
    return outs


# ----------------------------------------------------------------------------
# ConvBNReLU (models/modules.py:16-56) and DAFormer head (models/heads/daformer.py)
# ----------------------------------------------------------------------------
def conv_bn_act(p, x, stride=1, padding=0, dilation=1, groups=1, act="relu", bn_train=False, slope=0.1):
    x = F.conv2d(x, p("conv.weight"), p("conv.bias") if p.has("conv.bias") else None, stride, padding, dilation, groups)
    if p.has("bn.weight"):
        x = F.batch_norm(x, None if bn_train else p("bn.running_mean"), None if bn_train else p("bn.running_var"),
                         p("bn.weight"), p("bn.bias"), training=bn_train, eps=1e-5)
    if act == "relu":
        x = F.relu(x)
    elif act == "leaky":
        x = F.leaky_relu(x, slope)
    return x


def daformer_forward(sd, feats, prefix="", bn_train=False):
    """DAFormerHead.forward (:203-227) with dropout off; ASPP order [1x1, dw-sep d6, d12, d18]
    (:26-35,:52-62), bottleneck 3x3 (:102-108)."""
    p = SD(sd, prefix)
    os_size = feats[0].shape[2:]
    cs = []
    for i, f in enumerate(feats):
        n, c, h, w = f.shape
        e = p.sub("embed_layers.%d." % i)
        t = F.linear(f.flatten(2).transpose(1, 2), e("proj.weight"), e("proj.bias"))
        t = t.permute(0, 2, 1).reshape(n, -1, h, w)
        if t.shape[2:] != os_size:
            t = F.interpolate(t, size=os_size, mode="bilinear", align_corners=False)
        cs.append(t)
    x = torch.cat(cs, 1)
    a = p.sub("fuse_layer.aspp_modules.")
    C = x.shape[1]
    outs = [conv_bn_act(a.sub("0."), x, bn_train=bn_train)]
    for i, d in enumerate((6, 12, 18), start=1):
        t = conv_bn_act(a.sub("%d.depthwise_conv." % i), x, padding=d, dilation=d, groups=C, bn_train=bn_train)
        outs.append(conv_bn_act(a.sub("%d.pointwise_conv." % i), t, bn_train=bn_train))
    x = conv_bn_act(p.sub("fuse_layer.bottleneck."), torch.cat(outs, 1), padding=1, bn_train=bn_train)
    return F.conv2d(x, p("conv_seg.weight"), p("conv_seg.bias"))


# ----------------------------------------------------------------------------
# VGG-16 pyramid (models/backbones/vgg.py:108-149)
# ----------------------------------------------------------------------------
VGG16_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]


def vgg16_forward(sd, x, cut_points, prefix=""):
    """features[0:cut] for each cut in `cut_points` (indices into the nn.Sequential, e.g. [10, 17])."""
    p = SD(sd, prefix)
    outs, idx = [], 0
    for v in VGG16_CFG:
        if idx >= max(cut_points):
            break
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            idx += 1
        else:
            x = F.relu(F.conv2d(x, p("features.%d.weight" % idx), p("features.%d.bias" % idx), padding=1))
            idx += 2
        if idx in cut_points:
            outs.append(x)
    return outs


# ----------------------------------------------------------------------------
# UAWarpC head (models/heads/uawarpc.py, models/modules.py:395-561)
# ----------------------------------------------------------------------------
def flow_decoder(p, x):
    """OpticalFlowEstimatorResidualConnection.forward (modules.py:429-443), BN in eval mode."""
    c = lambda name, t, pad=1, act=None: conv_bn_act(p.sub(name + "."), t, padding=pad, act=act)
    x0 = c("conv_0", x)
    x2 = c("conv_2", c("conv_1", F.leaky_relu(x0, 0.1), act="leaky"))
    x2s = x2 + c("conv0_skip", x0, pad=0)
    x4 = c("conv_4", c("conv_3", F.leaky_relu(x2s, 0.1), act="leaky"))
    xo = F.leaky_relu(x4 + c("conv2_skip", x2s, pad=0), 0.1)
    return F.conv2d(xo, p("predict_mapping.weight"), p("predict_mapping.bias"), padding=1), xo


def refinement_module(p, x):
    """RefinementModule.forward (modules.py:446-477): dilations 1,2,4,8,16,1 then 3x3 -> 2."""
    for i, d in enumerate((1, 2, 4, 8, 16, 1)):
        x = conv_bn_act(p.sub("dc_convs.%d." % i), x, padding=d, dilation=d, act="leaky")
    return F.conv2d(x, p("dc_convs.6.weight"), p("dc_convs.6.bias"), padding=1)


def uncertainty_module(p, corr, feat, search, prev_u=None, prev_flow=None):
    """UncertaintyModule.forward (modules.py:534-561)."""
    b, _, h, w = corr.shape
    x = corr.permute(0, 2, 3, 1).reshape(b * h * w, 1, search, search)
    x = conv_bn_act(p.sub("conv_0."), x, act="leaky")
    if search == 16:
        x = F.max_pool2d(x, 2)
    x = conv_bn_act(p.sub("conv_2."), conv_bn_act(p.sub("conv_1."), x, act="leaky"), act="leaky")
    u = F.conv2d(x, p("predict_uncertainty.weight"), p("predict_uncertainty.bias"))
    u = u.reshape(b, h, w, -1).permute(0, 3, 1, 2)
    parts = (u, feat) if prev_u is None else (u, feat, prev_u, prev_flow)
    x = conv_bn_act(p.sub("pred_conv_0."), torch.cat(parts, 1), padding=1, act="leaky")
    x = conv_bn_act(p.sub("pred_conv_1."), x, padding=1, act="leaky")
    return F.conv2d(x, p("predict_uncertainty_final.weight"), p("predict_uncertainty_final.bias"), padding=1)


def mapping_to_flow(m):
    """unnormalise_and_convert_mapping_to_flow (helpers/matching_utils.py:77-103)."""
    B, _, H, W = m.shape
    xx = torch.arange(W, dtype=m.dtype).view(1, 1, 1, W)
    yy = torch.arange(H, dtype=m.dtype).view(1, 1, H, 1)
    fx = (m[:, 0:1] + 1) * (W - 1) / 2.0 - xx
    fy = (m[:, 1:2] + 1) * (H - 1) / 2.0 - yy
    return torch.cat((fx, fy), 1)


def _scale_xy(f, sx, sy):
    return torch.cat((f[:, 0:1] * sx, f[:, 1:2] * sy), 1)


def _up(x, size):
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


def uawarpc_forward(sd, trg, src, trg_256, src_256, out_size, prefix=""):
    """UAWarpCHead.forward (uawarpc.py:95-280) with estimate_uncertainty=True, both refinement
    modules, iterative_refinement=False.  trg/src = (1/4 feat, 1/8 feat) of the full-res images,
    *_256 = (1/8, 1/16) of the 256^2 images.  Returns 4 x (flow, log-variance)."""
    p = SD(sd, prefix)
    n = lambda t: F.normalize(t.float(), p=2, dim=1)
    c11, c12 = map(n, trg)
    c13, c14 = map(n, trg_256)
    c21, c22 = map(n, src)
    c23, c24 = map(n, src_256)
    h0, w0 = out_size
    # level 4 (global, 16x16)
    assert c14.shape[-2:] == (16, 16)
    corr4 = o_global_corr(c24, c14)
    est_map4, x4 = flow_decoder(p.sub("decoder4."), corr4)
    flow4_256 = _scale_xy(mapping_to_flow(est_map4), 256 / 16.0, 256 / 16.0)
    u4_256 = uncertainty_module(p.sub("estimate_uncertainty_components4."), corr4, x4, 16) + 2 * math.log(256 / 16.0)
    # level 3 (32x32, 256-px units)
    assert c13.shape[-2:] == (32, 32)
    up_flow4, up_u4 = _up(flow4_256, (32, 32)), _up(u4_256, (32, 32))
    warp3 = o_warp(c23, _scale_xy(up_flow4, 32 / 256.0, 32 / 256.0))
    corr3 = o_local_corr_layer(warp3, c13)
    res3, x3 = flow_decoder(p.sub("decoder3."), torch.cat((corr3, up_flow4, up_u4), 1))
    res3 = res3 + refinement_module(p.sub("refinement_module_adaptive."), x3)
    flow3 = res3 + up_flow4
    u3 = uncertainty_module(p.sub("estimate_uncertainty_components3."), corr3, x3, 9, up_u4, up_flow4)
    flow3 = _scale_xy(flow3, w0 / 256.0, h0 / 256.0)
    diag = math.sqrt(h0 ** 2 + w0 ** 2) / math.sqrt(256 ** 2 + 256 ** 2)
    u3 = u3 + 2 * math.log(diag)
    # level 2 (1/8)
    h2, w2 = c12.shape[-2:]
    up_flow3, up_u3 = _up(flow3, (h2, w2)), _up(u3, (h2, w2))
    warp2 = o_warp(c22, _scale_xy(up_flow3, w2 / float(w0), h2 / float(h0)))
    corr2 = o_local_corr_layer(warp2, c12)
    res2, x2 = flow_decoder(p.sub("decoder2."), torch.cat((corr2, up_flow3, up_u3), 1))
    flow2 = res2 + up_flow3
    u2 = uncertainty_module(p.sub("estimate_uncertainty_components2."), corr2, x2, 9, up_u3, up_flow3)
    # level 1 (1/4)
    h1, w1 = c11.shape[-2:]
    up_flow2, up_u2 = _up(flow2, (h1, w1)), _up(u2, (h1, w1))
    up_feat2 = F.conv2d(_up(x2, (h1, w1)), p("reduce.weight"), p("reduce.bias"))
    warp1 = o_warp(c21, _scale_xy(up_flow2, w1 / float(w0), h1 / float(h0)))
    corr1 = o_local_corr_layer(warp1, c11)
    res1, x1 = flow_decoder(p.sub("decoder1."), torch.cat((corr1, up_flow2, up_feat2, up_u2), 1))
    res1 = res1 + refinement_module(p.sub("refinement_module_finest."), x1)
    flow1 = res1 + up_flow2
    u1 = uncertainty_module(p.sub("estimate_uncertainty_components1."), corr1, x1, 9, up_u2, up_flow2)
    flow4 = _scale_xy(flow4_256, w0 / 256.0, h0 / 256.0)
    u4 = u4_256 + 2 * math.log(diag)
    return (flow4, u4), (flow3, u3), (flow2, u2), (flow1, u1)


def alignment_forward(vgg_sd, head_sd, images_i, images_j, vgg_prefix="", head_prefix=""):
    """AlignmentModel.forward (models/alignment_model.py:55-79) == the flow part of
    DomainAdaptationSegmentationModel.align (segmentation_model.py:493-520).
    Returns (flow i->j at image size, log-variance at image size)."""
    b, _, h, w = images_i.shape
    i256 = F.interpolate(images_i, size=(256, 256), mode="area")
    j256 = F.interpolate(images_j, size=(256, 256), mode="area")
    full = vgg16_forward(vgg_sd, torch.cat([images_j, images_i]), [10, 17], vgg_prefix)
    low = vgg16_forward(vgg_sd, torch.cat([j256, i256]), [17, 24], vgg_prefix)
    pyr_j, pyr_i = zip(*[torch.split(l, [b, b]) for l in full])
    pyr_j256, pyr_i256 = zip(*[torch.split(l, [b, b]) for l in low])
    flow, uncert = uawarpc_forward(head_sd, pyr_i, pyr_j, pyr_i256, pyr_j256, (h, w), head_prefix)[-1]
    return _up(flow, (h, w)), _up(uncert, (h, w))


def align(vgg_sd, head_sd, logits_ref, images_ref, images_trg, vgg_prefix="", head_prefix=""):
    """DomainAdaptationSegmentationModel.align (segmentation_model.py:493-523)."""
    flow, uncert = alignment_forward(vgg_sd, head_sd, images_trg, images_ref, vgg_prefix, head_prefix)
    warped, mask = o_warp(logits_ref, flow, return_mask=True)
    return warped, mask, o_cert(uncert), flow, uncert


# ----------------------------------------------------------------------------
# segmentation network + the target branch of training_step
# ----------------------------------------------------------------------------
def segmentor_logits(sd, images, backbone_prefix, head_prefix, model_type="mit_b5", bn_train=False):
    """head(backbone(x)) + bilinear upsample to the image size (segmentation_model.py:157-168)."""
    feats = mit_forward(sd, images, model_type, backbone_prefix)
    logits = daformer_forward(sd, feats, head_prefix, bn_train)
    return F.interpolate(logits, images.shape[-2:], mode="bilinear", align_corners=False), feats


def refign_target_branch(sd, images_trg, images_ref, model_type="mit_b5", gamma=0.25, bn_train=True):
    """The no-grad Refign branch of training_step (segmentation_model.py:201-214) + the torch.max
    of get_dacs_mix (:551): teacher logits on cat(trg, ref), align, refine, pseudo-label."""
    b = images_trg.shape[0]
    m_logits, _ = segmentor_logits(sd, torch.cat((images_trg, images_ref)), "m_backbone.", "m_head.", model_type, bn_train)
    lt, lr = torch.split(m_logits, [b, b], dim=0)
    warped, mask, certs, flow, uncert = align(sd, sd, lr, images_ref, images_trg, "alignment_backbone.", "alignment_head.")
    probs, label, maxprob, trust = o_refine(lt, warped, mask, certs=certs, gamma=gamma)
    return dict(logits_trg=lt, logits_ref=lr, warped=warped, mask=mask, certs=certs, flow=flow, uncert=uncert,
                probs=probs, label=label, maxprob=maxprob, trust=trust)


def pixel_weighted_ce(logits, target, pixel_weight=None, ignore_index=255):
    """PixelWeightedCrossEntropyLoss.forward (models/losses.py:16-22)."""
    loss = F.cross_entropy(logits, target, ignore_index=ignore_index, reduction="none")
    if pixel_weight is not None:
        loss = loss * pixel_weight
    return loss.mean()


def ema_update(ema_sd, live_sd, step, ema_momentum=0.999):
    """update_momentum_encoder (segmentation_model.py:680-689) over matching keys."""
    m = min(1.0 - 1 / (float(step) + 1.0), ema_momentum)
    return {k: ema_sd[k] * m + live_sd[k] * (1.0 - m) for k in ema_sd}
