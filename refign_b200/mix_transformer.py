"""MiT / SegFormer encoder with the reference's interface and ``state_dict`` keys
(reference: models/backbones/mix_transformer.py) on a channels-last token pipeline.

B200-first differences to the reference implementation:
  * tokens stay [B, N, C] (channels-last) end to end: the 3x3 depthwise conv of the Mix-FFN and the
    spatial-reduction conv read the token tensor through a zero-copy NHWC view instead of two
    NLC<->NCHW transpose copies per block (reference :96-103, :143-149);
  * the attention core softmax(q k^T * scale) v runs in the fused sm_100a kernel
    (refign_b200.ops.sr_attention): the [B, heads, N, N_kv] matrix the reference materialises and
    saves for backward (:156-160) never exists in HBM;
  * stage-1 OverlapPatchEmbed (7x7/s4 conv with 3 input channels + LayerNorm) is one kernel on the
    no-grad paths (EMA teacher, ImageNet copy).
Linear layers are plain library GEMMs (cuBLAS through torch).
"""
import math
import os
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .modules import DropPath


def _nhwc_view(x, H, W):
    """[B, H*W, C] tokens -> logical NCHW tensor with channels-last strides (no copy)."""
    B, N, C = x.shape
    return x.view(B, H, W, C).permute(0, 3, 1, 2)


def _tokens(x):
    """NCHW tensor (any strides) -> [B, H*W, C] tokens; zero-copy when x is channels-last."""
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B, H * W, C)


class DWConv(nn.Module):
    """3x3 depthwise conv over the token grid (reference :556-568)."""

    def __init__(self, dim=768):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)

    def forward(self, x, H, W):
        return _tokens(self.dwconv(_nhwc_view(x, H, W)))


class Mlp(nn.Module):
    """Mix-FFN: fc1 -> dwconv 3x3 -> GELU (exact erf) -> fc2 (reference :79-103)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.dwconv = DWConv(hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x, H, W):
        x = ops.linear(x, self.fc1.weight, self.fc1.bias)
        x = ops.dwconv3x3_gelu(x, H, W, self.dwconv.dwconv.weight, self.dwconv.dwconv.bias) \
            if ops.FUSED_DWCONV and isinstance(self.act, nn.GELU) else self.act(self.dwconv(x, H, W))
        x = self.drop(x)
        x = ops.linear(x, self.fc2.weight, self.fc2.bias)
        return self.drop(x)


class Attention(nn.Module):
    """Spatial-reduction attention (reference :106-164), head_dim 64 in every MiT variant."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., sr_ratio=1):
        super().__init__()
        assert dim % num_heads == 0
        self.dim = dim
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.sr_ratio = sr_ratio
        if sr_ratio > 1:
            self.sr = nn.Conv2d(dim, dim, kernel_size=sr_ratio, stride=sr_ratio)
            self.sr.weight._rf_sr_weight = True     # runtime.FlatParams stores it channels-last (the patch-GEMM layout)
            self.norm = nn.LayerNorm(dim)

    def forward(self, x, H, W):
        B, N, C = x.shape
        h, d = self.num_heads, C // self.num_heads
        q = ops.linear(x, self.q.weight, self.q.bias)        # [B, N, h*d]   (read strided per head)
        if self.sr_ratio > 1:
            x_ = ops.sr_conv(x, H, W, self.sr)
            x_ = ops.layer_norm(x_, self.norm)
        else:
            x_ = x
        kv = ops.linear(x_, self.kv.weight, self.kv.bias)    # [B, M, 2*h*d]: k = [..., :C], v = [..., C:]
        if self.attn_drop.p > 0 and self.training:
            raise NotImplementedError("attention dropout is not supported by the fused kernel "
                                      "(attn_drop_rate is 0 in every Refign config)")
        o = ops.sr_attention(q, kv, h, self.scale)           # [B, N, h*d]
        return self.proj_drop(ops.linear(o, self.proj.weight, self.proj.bias))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, sr_ratio=1):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop, sr_ratio=sr_ratio)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def _path_scale(self, x):
        """Per-sample drop-path factor mask / keep_prob as a [B] tensor, or None (reference
        models/modules.py:587-596)."""
        dp = self.drop_path
        if not isinstance(dp, DropPath) or not dp.drop_prob or not dp.training:
            return None
        keep = 1.0 - dp.drop_prob
        return torch.empty(x.shape[0], dtype=torch.float32, device=x.device).bernoulli_(keep) / keep

    def forward(self, x, H, W, carry=None, path_scales=None):
        """x: residual stream [B,N,C]; ``carry`` = (branch, scale) of the previous block's still
        un-added MLP branch.  The residual adds are fused into the LayerNorm that follows them
        (refign_b200.ops.add_layer_norm), so this returns (x, carry) with the last add pending:
            x = x + dp(attn(norm1(x))); x = x + dp(mlp(norm2(x)))          (reference :203-207)
        ``path_scales``: the two per-sample drop-path factors of this block, pre-drawn for the whole backbone in one
        launch by MixVisionTransformer.forward_features (else drawn here, two small launches per branch)."""
        if carry is None:
            h = ops.layer_norm(x, self.norm1)
        else:
            x, h = ops.add_layer_norm(x, carry[0], carry[1], self.norm1)
        a = self.attn(h, H, W)
        s_attn, s_mlp = path_scales if path_scales is not None else (self._path_scale(x), self._path_scale(x))
        x, h = ops.add_layer_norm(x, a, s_attn, self.norm2)
        m = self.mlp(h, H, W)
        return x, (m, s_mlp)


class OverlapPatchEmbed(nn.Module):
    """Overlapping patch embedding: strided conv + LayerNorm(eps 1e-5) (reference :210-242)."""

    def __init__(self, img_size=224, patch_size=7, stride=4, in_chans=3, embed_dim=768):
        super().__init__()
        self.patch_size = (patch_size, patch_size)
        self.stride = stride
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=stride,
                              padding=(patch_size // 2, patch_size // 2))
        self.norm = nn.LayerNorm(embed_dim)

    def forward(self, x):
        fused = (ops.FUSED_PATCH_EMBED and self.proj.in_channels == 3 and self.patch_size == (7, 7)
                 and self.stride == 4 and self.proj.out_channels in (32, 64) and self.proj.bias is not None)
        if fused:
            return ops.patch_embed_ln(x, self.proj.weight, self.proj.bias, self.norm.weight, self.norm.bias,
                                      self.norm.eps)
        y = self.proj(x.contiguous(memory_format=torch.channels_last) if x.is_cuda and x.shape[1] > 3 else x)
        _, _, H, W = y.shape
        # the residual stream starts here and stays fp32 (as under the reference's AMP, where LayerNorm
        # autocasts to fp32)
        return ops.layer_norm(_tokens(y), self.norm, out_dtype=torch.float32), H, W


class MixVisionTransformer(nn.Module):
    _ln6 = partial(nn.LayerNorm, eps=1e-6)
    arch_settings = {
        name: dict(embed_dims=dims, num_heads=heads, mlp_ratios=[4, 4, 4, 4], qkv_bias=True, depths=depths,
                   sr_ratios=[8, 4, 2, 1])
        for name, dims, heads, depths in (
            ('mit_b0', [32, 64, 160, 256], [1, 2, 5, 8], [2, 2, 2, 2]),
            ('mit_b1', [64, 128, 320, 512], [1, 2, 5, 8], [2, 2, 2, 2]),
            ('mit_b2', [64, 128, 320, 512], [1, 2, 5, 8], [3, 4, 6, 3]),
            ('mit_b3', [64, 128, 320, 512], [1, 2, 5, 8], [3, 4, 18, 3]),
            ('mit_b4', [64, 128, 320, 512], [1, 2, 5, 8], [3, 8, 27, 3]),
            ('mit_b5', [64, 128, 320, 512], [1, 2, 5, 8], [3, 6, 40, 3]),
        )
    }

    def __init__(self, model_type, pretrained=None, img_size=224, in_chans=3, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0.1, freeze_patch_embed=False):
        super().__init__()
        cfg = self.arch_settings[model_type]
        dims, heads, ratios, depths, srs = (cfg['embed_dims'], cfg['num_heads'], cfg['mlp_ratios'], cfg['depths'],
                                            cfg['sr_ratios'])
        self.model_type = model_type
        self.depths = depths
        self.embed_dims = dims
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths))]
        cur = 0
        for s in range(4):
            setattr(self, 'patch_embed%d' % (s + 1), OverlapPatchEmbed(
                img_size=img_size if s == 0 else img_size // (2 ** (s + 1)), patch_size=7 if s == 0 else 3,
                stride=4 if s == 0 else 2, in_chans=in_chans if s == 0 else dims[s - 1], embed_dim=dims[s]))
        for s in range(4):
            setattr(self, 'block%d' % (s + 1), nn.ModuleList([
                Block(dim=dims[s], num_heads=heads[s], mlp_ratio=ratios[s], qkv_bias=cfg['qkv_bias'],
                      qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[cur + i],
                      norm_layer=self._ln6, sr_ratio=srs[s]) for i in range(depths[s])]))
            setattr(self, 'norm%d' % (s + 1), self._ln6(dims[s]))
            cur += depths[s]
        if freeze_patch_embed:
            self.freeze_patch_emb()
        self.init_weights(pretrained=pretrained)

    # ---- weights ----------------------------------------------------------------------------------
    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Conv2d):
            fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels // m.groups
            m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
            if m.bias is not None:
                m.bias.data.zero_()

    def init_weights(self, pretrained=None):
        if pretrained is None:
            self.apply(self._init_weights)
            return
        path = resolve_checkpoint(pretrained, self.model_type)
        ckpt = torch.load(path, map_location='cpu')
        sd = ckpt.get('state_dict', ckpt.get('model', ckpt)) if isinstance(ckpt, dict) else ckpt
        if any(k.startswith('backbone.') for k in sd):
            sd = {k[len('backbone.'):]: v for k, v in sd.items() if k.startswith('backbone.')}
        sd = {k: v for k, v in sd.items() if not k.startswith('head.')}
        self.load_state_dict(sd, strict=True)

    def reset_drop_path(self, drop_path_rate):
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(self.depths))]
        cur = 0
        for s in range(4):
            for i, blk in enumerate(getattr(self, 'block%d' % (s + 1))):
                if isinstance(blk.drop_path, DropPath):
                    blk.drop_path.drop_prob = dpr[cur + i]
            cur += self.depths[s]

    def freeze_patch_emb(self):
        # reference quirk kept: this sets an attribute and freezes nothing
        self.patch_embed1.requires_grad = False

    # ---- forward ----------------------------------------------------------------------------------
    def _draw_path_scales(self, B, device):
        """All drop-path factors of one forward, mask / keep_prob per (block, branch, sample), as ONE [2 * blocks, B]
        tensor drawn with three launches (the per-branch draw of reference models/modules.py:587-596 is two small launches
        x 104 branches x three training forwards per step).  None when nothing is dropped."""
        if not self.training:
            return None
        probs = []
        for s in range(4):
            for blk in getattr(self, 'block%d' % (s + 1)):
                dp = blk.drop_path
                probs.append(float(dp.drop_prob) if isinstance(dp, DropPath) and dp.drop_prob and dp.training else 0.0)
        if not any(probs):
            return None
        key = (tuple(probs), str(device))
        cached = getattr(self, '_rf_keep', None)
        if cached is None or cached[0] != key:
            keep = torch.tensor([1.0 - p for p in probs for _ in (0, 1)], dtype=torch.float32, device=device).unsqueeze(1)
            cached = (key, keep)
            self._rf_keep = cached
        keep = cached[1]
        return (torch.rand(keep.shape[0], B, device=device) < keep).to(torch.float32) / keep

    def forward_features(self, x):
        outs = []
        B = x.shape[0]
        scales = self._draw_path_scales(B, x.device) if x.is_cuda else None
        bi = 0
        for s in range(4):
            x, H, W = getattr(self, 'patch_embed%d' % (s + 1))(x)
            carry = None
            for blk in getattr(self, 'block%d' % (s + 1)):
                x, carry = blk(x, H, W, carry, None if scales is None else (scales[2 * bi], scales[2 * bi + 1]))
                bi += 1
            norm = getattr(self, 'norm%d' % (s + 1))
            # stage outputs keep the residual stream's dtype (fp32, like LayerNorm under the reference's AMP):
            # they feed the feature-distance loss as well as the decode head
            if carry is None:
                x = ops.layer_norm(x, norm, out_dtype=x.dtype)
            else:
                _, x = ops.add_layer_norm(x, carry[0], carry[1], norm, out_dtype=x.dtype)
            # stage output: logical NCHW, physically channels-last (zero-copy view of the tokens);
            # consumers that need plain NCHW call .contiguous()
            x = x.view(B, H, W, -1).permute(0, 3, 1, 2)
            outs.append(x)
        return outs

    def forward(self, x):
        return self.forward_features(x)


def resolve_checkpoint(pretrained, model_type=None):
    """Local path lookup identical in spirit to the reference (:449-462); there is no network here,
    so an unresolvable name raises instead of downloading."""
    cands = [pretrained, os.path.join(os.environ.get('TORCH_HOME', ''), 'hub', pretrained)]
    if pretrained in ('imagenet', 'cityscapes') and model_type:
        cands += [os.path.join('pretrained_models', '%s.pth' % model_type),
                  os.path.join(os.environ.get('TORCH_HOME', ''), 'hub', 'checkpoints', '%s.pth' % model_type)]
    for c in cands:
        if c and os.path.isfile(c):
            return c
    raise FileNotFoundError("checkpoint '%s' not found locally (looked in %s); downloads are not available"
                            % (pretrained, cands))
