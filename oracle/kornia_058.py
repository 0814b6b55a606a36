"""CPU restatement of the kornia 0.5.8 operators the reference's DACS strong transform calls -- TEST INFRASTRUCTURE.

The reference (helpers/dacs_transforms.py:42-78, called per image from models/segmentation_model.py:561-570) uses the
third-party dependency ``kornia==0.5.8`` (requirements.txt:7), which is NOT vendored in /root/reference and not
installable here (no network, not in the wheelhouse).  PARITY UNPINNED against a live kornia: what follows restates the
published 0.5.8 algorithms (kornia/augmentation/augmentation.py ``ColorJitter.apply_transform``,
kornia/enhance/adjust.py ``adjust_brightness / adjust_contrast / adjust_saturation / adjust_hue``,
kornia/color/hsv.py ``rgb_to_hsv / hsv_to_rgb``, kornia/filters/gaussian.py + kernels.py ``GaussianBlur2d``) as plain
torch on CPU, anchored on the reference's own call sites:
  * ColorJitter(brightness=s, contrast=s, saturation=s, hue=s), p = 1:  the four adjustments are applied in a random
    order (``torch.randperm(4)``);  brightness is ADDITIVE (``adjust_brightness(img, factor - 1)``, factor ~ U(max(0, 1-s),
    min(2, 1+s))), contrast MULTIPLICATIVE (``img * factor``, factor ~ U(max(0, 1-s), 1+s)), both clamped to [0, 1];
    saturation scales S of HSV (factor ~ U(max(0, 1-s), 1+s), clamp [0, 1]); hue adds ``factor * 2 pi`` to H
    (factor ~ U(-s, s), |s| <= 0.5) with ``torch.fmod(h + shift, 2 pi)``.
  * GaussianBlur2d(kernel_size=(ky, kx), sigma=(sigma, sigma)), border_type='reflect': normalised 1-D gaussians
    ``exp(-x^2 / (2 sigma^2))`` on x = arange(k) - k // 2 (+ 0.5 for even k), outer product, zero-phase correlation.
The random draws themselves are outside the restatement (kornia draws them from torch's global RNG; the product draws
the same distributions from Python's ``random``): functions here are deterministic in their explicit parameters.
"""
import math

import torch
import torch.nn.functional as F

_EPS = 1e-6   # kornia.color.hsv.rgb_to_hsv default


def rgb_to_hsv(image):
    """[*, 3, H, W] in [0, 1] -> (h in [0, 2 pi), s, v);  kornia/color/hsv.py (0.5.8)."""
    max_rgb, argmax_rgb = image.max(-3)
    min_rgb = image.min(-3)[0]
    deltac = max_rgb - min_rgb
    v = max_rgb
    s = deltac / (max_rgb + _EPS)
    deltac = torch.where(deltac == 0, torch.ones_like(deltac), deltac)
    rc, gc, bc = torch.unbind(max_rgb.unsqueeze(-3) - image, dim=-3)
    h1 = bc - gc
    h2 = (rc - bc) + 2.0 * deltac
    h3 = (gc - rc) + 4.0 * deltac
    h = torch.stack((h1, h2, h3), dim=-3) / deltac.unsqueeze(-3)
    h = torch.gather(h, dim=-3, index=argmax_rgb.unsqueeze(-3)).squeeze(-3)
    h = (h / 6.0) % 1.0
    h = 2.0 * math.pi * h
    return torch.stack((h, s, v), dim=-3)


def hsv_to_rgb(image):
    h = image[..., 0, :, :] / (2 * math.pi)
    s = image[..., 1, :, :]
    v = image[..., 2, :, :]
    hi = torch.floor(h * 6) % 6
    f = ((h * 6) % 6) - hi
    one = torch.tensor(1.0)
    p = v * (one - s)
    q = v * (one - f * s)
    t = v * (one - (one - f) * s)
    hi = hi.long()
    indices = torch.stack([hi, hi + 6, hi + 12], dim=-3)
    out = torch.stack((v, q, p, p, t, v, t, v, v, q, p, p, p, p, t, v, v, q), dim=-3)
    return torch.gather(out, -3, indices)


def adjust_brightness(x, factor):
    return (x + factor).clamp(0.0, 1.0)


def adjust_contrast(x, factor):
    return (x * factor).clamp(0.0, 1.0)


def adjust_saturation(x, factor):
    h, s, v = torch.chunk(rgb_to_hsv(x), chunks=3, dim=-3)
    return hsv_to_rgb(torch.cat([h, torch.clamp(s * factor, min=0, max=1), v], dim=-3))


def adjust_hue(x, factor):
    h, s, v = torch.chunk(rgb_to_hsv(x), chunks=3, dim=-3)
    return hsv_to_rgb(torch.cat([torch.fmod(h + factor, 2 * math.pi), s, v], dim=-3))


def color_jitter(x, order, brightness, contrast, saturation, hue):
    """ColorJitter.apply_transform for ONE image batch x [B, 3, H, W] in [0, 1] with explicit parameters:
    ``order`` a permutation of (0, 1, 2, 3) = (brightness, contrast, saturation, hue)."""
    ops = (lambda im: adjust_brightness(im, brightness - 1.0), lambda im: adjust_contrast(im, contrast),
           lambda im: adjust_saturation(im, saturation), lambda im: adjust_hue(im, hue * 2 * math.pi))
    for idx in order:
        x = ops[int(idx)](x)
    return x


def gaussian_kernel1d(ksize, sigma):
    x = torch.arange(ksize, dtype=torch.float32) - ksize // 2
    if ksize % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * sigma ** 2))
    return g / g.sum()


def gaussian_blur2d(x, kernel_size, sigma):
    """kornia.filters.GaussianBlur2d(kernel_size, (sigma, sigma), border_type='reflect') on [B, C, H, W]."""
    ky, kx = kernel_size
    k2 = torch.matmul(gaussian_kernel1d(ky, sigma).unsqueeze(-1), gaussian_kernel1d(kx, sigma).unsqueeze(-1).t())
    C = x.shape[1]
    xp = F.pad(x, (kx // 2, kx // 2, ky // 2, ky // 2), mode='reflect')
    return F.conv2d(xp, k2.view(1, 1, ky, kx).repeat(C, 1, 1, 1), groups=C)


_MEAN = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
_STD = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)


def dacs_strong_transform(img_src, img_trg, gt_src, pseudo_label, pseudo_weight, mask, jitter=None, blur=None):
    """helpers/dacs_transforms.py:14-27 for one image pair with explicit random parameters.
    mask [1, H, W] (1 = take the source pixel); jitter = None | (order, brightness, contrast, saturation, hue);
    blur = None | (ky, kx, sigma).  Returns (mixed image [1,3,H,W], mixed label [1,H,W], mixed weight [1,H,W])."""
    m = mask.to(img_src.dtype)
    img = (m * img_src + (1 - m) * img_trg).unsqueeze(0)
    ml = mask.to(gt_src.dtype)
    lbl = ml * gt_src + (1 - ml) * pseudo_label
    w = m * torch.ones_like(pseudo_weight) + (1 - m) * pseudo_weight
    if jitter is not None:
        img = (color_jitter(img * _STD + _MEAN, *jitter) - _MEAN) / _STD
    if blur is not None:
        img = gaussian_blur2d(img, (blur[0], blur[1]), blur[2])
    return img, lbl, w
