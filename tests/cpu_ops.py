"""TEST-ONLY: run the host-side modules of refign_b200 on CPU by routing the operator layer
(refign_b200.ops) to the CPU oracle.  This exercises the host logic (module wiring, state_dict
compatibility, scale bookkeeping, runtime) without a GPU; the kernels themselves are tested by the
``-m gpu`` tests.  Never used by the product."""
import contextlib

import torch
import torch.nn.functional as F

import oracle
from refign_b200 import ops


def _refine(lt, lr, warp_mask=None, certs=None, logvar=None, gamma=0.25, disable_M=False, disable_P=False,
            static_classes=ops.STATIC_LARGE_CLASSES, want_label=True):
    return oracle.refine(lt, lr, warp_mask, certs=certs, logvar=logvar, gamma=gamma, disable_M=disable_M,
                         disable_P=disable_P)


def _ema(ema, live, m):
    ema.mul_(float(torch.tensor(m, dtype=torch.float32))).add_(live * float(torch.tensor(1.0 - m, dtype=torch.float32)))
    return ema


def _adamw(param, grad, exp_avg, exp_avg_sq, seg_end, seg_lr, seg_wd, beta1, beta2, eps, step, grad_scale=1.0):
    start = 0
    for end, lr, wd in zip(seg_end, seg_lr, seg_wd):
        p, g = param[start:end], grad[start:end] * grad_scale
        m, v = exp_avg[start:end], exp_avg_sq[start:end]
        p.mul_(1 - lr * wd)
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        p.addcdiv_(m, (v.sqrt() / (bc2 ** 0.5)).add_(eps), value=-lr / bc1)
        start = end
    return param


def _dwconv_gelu(x, H, W, weight, bias):
    B, N, C = x.shape
    y = F.conv2d(x.transpose(1, 2).reshape(B, C, H, W), weight, bias, padding=1, groups=C)
    return F.gelu(y).flatten(2).transpose(1, 2)


def _dwconv_nhwc(x, weight, bias=None, dilation=1):
    return F.conv2d(x, weight, bias, padding=dilation, dilation=dilation, groups=x.shape[1])


def _layer_norm(x, norm, out_dtype=None):
    return F.layer_norm(x, (x.shape[-1],), norm.weight, norm.bias, norm.eps)


def _add_layer_norm(x, branch, scale, norm, out_dtype=None):
    xn = x + (branch if scale is None else branch * scale.view(-1, 1, 1))
    return xn, F.layer_norm(xn, (xn.shape[-1],), norm.weight, norm.bias, norm.eps)


def _ema_dev(ema, live, hyper):
    return ema.mul_(float(hyper[10])).add_(live * float(hyper[11]))


def _adamw_dev(param, grad, exp_avg, exp_avg_sq, seg_end, seg_wd, beta1, beta2, eps, hyper, grad_scale=1.0):
    start = 0
    bc1, sq2 = float(hyper[8]), float(hyper[9])
    for i, (end, wd) in enumerate(zip(seg_end, seg_wd)):
        lr = float(hyper[i])
        p, g = param[start:end], grad[start:end] * grad_scale
        m, v = exp_avg[start:end], exp_avg_sq[start:end]
        p.mul_(1 - lr * wd)
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        p.addcdiv_(m, (v.sqrt() / sq2).add_(eps), value=-lr / bc1)
        start = end
    return param


def _patch_embed_ln(x, conv_w, conv_b, ln_w, ln_b, eps):
    y = F.conv2d(x, conv_w, conv_b, stride=4, padding=3)
    H, W = y.shape[2:]
    return F.layer_norm(y.flatten(2).transpose(1, 2), (y.shape[1],), ln_w, ln_b, eps), H, W


def _warp(x, flo, padding_mode="zeros", return_mask=False):
    """Oracle warp; when a gradient is needed (alignment training) the same sampling through torch's
    differentiable grid_sample (align_corners=True, zeros padding: what helpers/matching_utils.py:11-49 calls)."""
    if not (torch.is_grad_enabled() and (x.requires_grad or flo.requires_grad)):
        return oracle.warp(x, flo, return_mask=return_mask)
    B, C, H, W = x.shape
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, W).expand(B, H, W)
    ys = torch.arange(H, dtype=torch.float32).view(1, H, 1).expand(B, H, W)
    gx = 2.0 * (xs + flo[:, 0].float()) / max(W - 1, 1) - 1.0
    gy = 2.0 * (ys + flo[:, 1].float()) / max(H - 1, 1) - 1.0
    grid = torch.stack((gx, gy), -1)
    out = F.grid_sample(x.float(), grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    if return_mask:
        mask = (gx > -1) & (gy > -1) & (gx < 1) & (gy < 1)
        return out, mask
    return out


def _local_corr(s, t, P=9):
    """Oracle local correlation layer; differentiable shift-and-sum formulation when a gradient is needed."""
    if not (torch.is_grad_enabled() and (s.requires_grad or t.requires_grad)):
        return oracle.local_corr_layer(s, t, P)
    r = (P - 1) // 2
    B, C, H, W = s.shape
    sp = F.pad(s.float(), (r, r, r, r))
    planes = [(t.float() * sp[:, :, dy:dy + H, dx:dx + W]).sum(1) for dy in range(P) for dx in range(P)]
    corr = torch.stack(planes, 1)
    return F.normalize(F.relu(corr), p=2, dim=1)


def _sr_attention(q, kv, heads, scale):
    """The reference's materialising formulation (mix_transformer.py:150-160: q k^T * scale, softmax, @ v) on the
    packed q [B,N,h*d] / kv [B,M,2*h*d] layout -- the CPU stand-in for the fused CUDA kernels."""
    B, N, C = q.shape
    M = kv.shape[1]
    d = C // heads
    q4 = q.view(B, N, heads, d).transpose(1, 2)
    k4 = kv[..., :C].reshape(B, M, heads, d).transpose(1, 2)
    v4 = kv[..., C:].reshape(B, M, heads, d).transpose(1, 2)
    attn = torch.softmax((q4 @ k4.transpose(-2, -1)) * scale, dim=-1)
    return (attn @ v4).transpose(1, 2).reshape(B, N, C)


_PATCH = {
    "sr_attention": _sr_attention,
    "patch_embed_ln": _patch_embed_ln,
    "ema_update_dev_": _ema_dev,
    "adamw_step_dev_": _adamw_dev,
    "layer_norm": _layer_norm,
    "add_layer_norm": _add_layer_norm,
    "dwconv3x3_gelu": _dwconv_gelu,
    "dwconv3x3_nhwc": _dwconv_nhwc,
    "local_correlation_relu_l2norm": _local_corr,
    "global_correlation": lambda s, t, cyclic_consistency=True, normalise=True, use_tensor_cores=-1:
        oracle.global_corr(s, t),
    "warp": _warp,
    "estimate_probability_of_confidence_interval_of_mixture_density": lambda u, R=1.0: oracle.cert(u),
    "refine_fused": _refine,
    "ema_update_": _ema,
    "adamw_step_": _adamw,
}


@contextlib.contextmanager
def cpu_ops():
    saved = {k: getattr(ops, k) for k in _PATCH}
    try:
        for k, v in _PATCH.items():
            setattr(ops, k, v)
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
