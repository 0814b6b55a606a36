"""DACS class-mix + colour jitter + gaussian blur between ``refine`` and the student forward
(reference helpers/dacs_transforms.py, kornia-based).  Rebuilt without kornia and without
the host synchronisations of the reference (``torch.unique`` + ``classes.shape[0]``, dacs_transforms.py:84-86): the
class subset is drawn on the device, the random augmentation parameters on the host into a per-image parameter
block, and the whole transform of a batch is ONE fused kernel (+ two separable-blur passes) on the GPU
(refign_b200/csrc/dacs.cu through ops.dacs_mix) instead of a per-image Python loop of ~40 small launches.

Parity: the class-mix arithmetic (``one_mix``) is exact.  The jitter / blur follow kornia 0.5.8 (the reference's pinned
third-party dependency, absent here -- "parity unpinned" against a live kornia) as restated in oracle/kornia_058.py
from the published source; tests/test_dacs_gpu.py holds the kernel to that restatement, and this file's torch mirror
to it on the CPU.  The random streams themselves cannot match (kornia draws from torch's global RNG).
"""
import math
import random

import torch
import torch.nn.functional as F

_MEAN = (0.485, 0.456, 0.406)
_STD = (0.229, 0.224, 0.225)


def _stats(img):
    mean = img.new_tensor(_MEAN).view(1, 3, 1, 1)
    std = img.new_tensor(_STD).view(1, 3, 1, 1)
    return mean, std


def denorm(img):
    mean, std = _stats(img)
    return img * std + mean


def renorm(img):
    mean, std = _stats(img)
    return (img - mean) / std


def get_class_masks(labels, generator=None):
    """One mask per image selecting half (rounded up) of the classes present in the WHOLE batch
    (the reference's kept quirk, dacs_transforms.py:84-85), chosen uniformly at random per image.
    labels: [B,1,H,W] int64.  Returns a list of B tensors [1,1,H,W] (int64 0/1).  No host sync."""
    B = labels.shape[0]
    dev = labels.device
    present = torch.zeros(256, dtype=torch.bool, device=dev)
    present[labels.reshape(-1).clamp(0, 255)] = True
    n = present.sum()
    k = (n + n % 2) // 2
    scores = torch.rand(B, 256, device=dev, generator=generator)
    scores = torch.where(present.unsqueeze(0), scores, torch.full_like(scores, 2.0))
    rank = scores.argsort(dim=1).argsort(dim=1)          # rank of each class among the random scores
    chosen = (rank < k) & present.unsqueeze(0)           # [B,256]
    masks = torch.gather(chosen.long(), 1, labels.reshape(B, -1).clamp(0, 255)).view(labels.shape)
    return [m.unsqueeze(0) for m in masks]


def one_mix(mask, data=None, target=None):
    """mask*first + (1-mask)*second (dacs_transforms.py:101-112)."""
    if mask is None:
        return data, target
    if data is not None:
        m = mask[0].to(data.dtype)
        data = (m * data[0] + (1 - m) * data[1]).unsqueeze(0)
    if target is not None:
        m = mask[0].to(target.dtype)
        target = (m * target[0] + (1 - m) * target[1]).unsqueeze(0)
    return data, target


# ---- kornia 0.5.8 ColorJitter / GaussianBlur2d semantics (requirements.txt:7 of the reference) -----------------
# The random draws are made on the HOST into a small per-image parameter block (``draw_strong_params``); applying them
# is a deterministic function of that block -- on CUDA tensors ONE fused kernel (ops.dacs_mix: class mix + jitter,
# + the separable blur kernel), on CPU tensors the plain torch mirror below (same formulas; the host-side tests run it).
PARAM_STRIDE = 64
_P_JITTER, _P_ORDER, _P_BSHIFT, _P_CONTRAST, _P_SAT, _P_HSHIFT = 0, 1, 5, 6, 7, 8
_P_BLUR, _P_RY, _P_RX, _P_WY, _P_WX, _MAX_R = 9, 10, 11, 12, 29, 16
_TWO_PI = 2 * math.pi


def blur_kernel_size(n):
    """dacs_transforms.py:66-73: ~10 % of the image side, forced odd."""
    c = math.ceil(0.1 * n)
    return int(math.floor(c - 0.5 + c % 2))


def _half_kernel(ksize, sigma):
    """Centre + one side of kornia's normalised 1-D gaussian (filters/kernels.py gaussian()), truncated where it is
    < 1e-12 of the centre (8 sigma; the reference's 103-tap kernel at 1024 px has 21 non-negligible taps)."""
    r = min(ksize // 2, int(8 * sigma + 1), _MAX_R)
    w = [math.exp(-(x * x) / (2 * sigma * sigma)) for x in range(r + 1)]
    tot = w[0] + 2 * sum(w[1:])
    return r, [v / tot for v in w]


def draw_strong_params(B, H, W, color_jitter, s, p, blur, rng=random):
    """The random parameters of ``strong_transform`` for B images as a float32 [B, 64] CPU tensor.
    ``color_jitter`` / ``blur`` are the per-batch uniform draws of get_dacs_mix (segmentation_model.py:544-549); per image
    kornia's ColorJitter draws an order and four factors (brightness U(max(0,1-s), min(2,1+s)), contrast and
    saturation U(max(0,1-s), 1+s), hue U(-s, s) with s <= 0.5) and gaussian_blur a sigma ~ U(0.15, 1.15)."""
    P = torch.zeros(B, PARAM_STRIDE, dtype=torch.float32)
    jitter_on = bool(color_jitter > p)
    blur_on = bool(blur > 0.5)
    if isinstance(s, dict):
        sb, sc, ss, sh = (float(s.get(k, 0.0)) for k in ('brightness', 'contrast', 'saturation', 'hue'))
    else:
        sb = sc = ss = sh = float(s)
    sh = min(sh, 0.5)
    for b in range(B):
        if jitter_on:
            order = [0, 1, 2, 3]
            rng.shuffle(order)
            P[b, _P_JITTER] = 1.0
            for t in range(4):
                P[b, _P_ORDER + t] = float(order[t])
            P[b, _P_BSHIFT] = rng.uniform(max(0.0, 1 - sb), min(2.0, 1 + sb)) - 1.0
            P[b, _P_CONTRAST] = rng.uniform(max(0.0, 1 - sc), 1 + sc)
            P[b, _P_SAT] = rng.uniform(max(0.0, 1 - ss), 1 + ss)
            P[b, _P_HSHIFT] = rng.uniform(-sh, sh) * _TWO_PI
        if blur_on:
            sigma = rng.uniform(0.15, 1.15)
            ry, wy = _half_kernel(blur_kernel_size(H), sigma)
            rx, wx = _half_kernel(blur_kernel_size(W), sigma)
            P[b, _P_BLUR], P[b, _P_RY], P[b, _P_RX] = 1.0, float(ry), float(rx)
            P[b, _P_WY:_P_WY + ry + 1] = torch.tensor(wy)
            P[b, _P_WX:_P_WX + rx + 1] = torch.tensor(wx)
    return P


def _rgb_to_hsv(x):
    """kornia/color/hsv.py (0.5.8): h in [0, 2 pi), eps = 1e-6."""
    mx, arg = x.max(-3)
    mn = x.min(-3)[0]
    dc = mx - mn
    s = dc / (mx + 1e-6)
    dc = torch.where(dc == 0, torch.ones_like(dc), dc)
    rc, gc, bc = torch.unbind(mx.unsqueeze(-3) - x, dim=-3)
    h = torch.stack((bc - gc, (rc - bc) + 2.0 * dc, (gc - rc) + 4.0 * dc), dim=-3) / dc.unsqueeze(-3)
    h = torch.gather(h, -3, arg.unsqueeze(-3)).squeeze(-3)
    return _TWO_PI * ((h / 6.0) % 1.0), s, mx


def _hsv_to_rgb(h, s, v):
    h = h / _TWO_PI
    hi = torch.floor(h * 6) % 6
    f = ((h * 6) % 6) - hi
    p, q, t = v * (1 - s), v * (1 - f * s), v * (1 - (1 - f) * s)
    hi = hi.long()
    out = torch.stack((v, q, p, p, t, v, t, v, v, q, p, p, p, p, t, v, v, q), dim=-3)
    return torch.gather(out, -3, torch.stack([hi, hi + 6, hi + 12], dim=-3))


def apply_color_jitter(data, prow):
    """ColorJitter.apply_transform of kornia 0.5.8 on a NORMALISED image batch [*, 3, H, W] with the parameters of one
    row of ``draw_strong_params`` (denorm -> four adjustments in the drawn order -> renorm, dacs_transforms.py:52-58)."""
    if float(prow[_P_JITTER]) == 0.0:
        return data
    x = denorm(data)
    for t in range(4):
        op = int(prow[_P_ORDER + t])
        if op == 0:
            x = (x + float(prow[_P_BSHIFT])).clamp(0.0, 1.0)
        elif op == 1:
            x = (x * float(prow[_P_CONTRAST])).clamp(0.0, 1.0)
        else:
            h, sat, v = _rgb_to_hsv(x)
            if op == 2:
                sat = (sat * float(prow[_P_SAT])).clamp(0.0, 1.0)
            else:
                h = torch.fmod(h + float(prow[_P_HSHIFT]), _TWO_PI)
            x = _hsv_to_rgb(h, sat, v)
    return renorm(x)


def apply_gaussian_blur(data, prow):
    """GaussianBlur2d (reflect border) with the half kernels of one parameter row, as two 1-D correlations."""
    if float(prow[_P_BLUR]) == 0.0:
        return data
    ry, rx = int(prow[_P_RY]), int(prow[_P_RX])
    wy = torch.cat((prow[_P_WY + 1:_P_WY + ry + 1].flip(0), prow[_P_WY:_P_WY + ry + 1])).to(data)
    wx = torch.cat((prow[_P_WX + 1:_P_WX + rx + 1].flip(0), prow[_P_WX:_P_WX + rx + 1])).to(data)
    C = data.shape[1]
    x = F.pad(data, (rx, rx, ry, ry), mode='reflect')
    x = F.conv2d(x, wx.view(1, 1, 1, -1).repeat(C, 1, 1, 1), groups=C)
    return F.conv2d(x, wy.view(1, 1, -1, 1).repeat(C, 1, 1, 1), groups=C)


def strong_transform(param, data=None, target=None):
    """helpers/dacs_transforms.py:14-27 for one image: ``param['mix']`` the class mask, ``param['row']`` one row of
    ``draw_strong_params`` (None: no jitter / blur)."""
    assert data is not None or target is not None
    data, target = one_mix(mask=param['mix'], data=data, target=target)
    row = param.get('row')
    if data is not None and data.shape[1] == 3 and row is not None:
        data = apply_gaussian_blur(apply_color_jitter(data, row), row)
    return data, target
