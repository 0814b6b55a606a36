"""DACS class-mix + colour jitter + gaussian blur between ``refine`` and the student forward
(reference helpers/dacs_transforms.py, kornia-based).  Kept in plain torch -- it is outside the
named kernels -- but rebuilt without kornia and without the host
synchronisations of the reference (``torch.unique`` + ``classes.shape[0]``, dacs_transforms.py:84-86):
the class subset is drawn on the device.

PARITY UNPINNED for the random augmentations (kornia 0.5.8 is not installable here and the step is
random by construction); the class-mix arithmetic (``one_mix``) is exact.
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


def _gray(x):
    return (0.299 * x[:, 0:1] + 0.587 * x[:, 1:2] + 0.114 * x[:, 2:3])


def color_jitter(color_jitter, data=None, target=None, s=.25, p=.2):
    """Brightness / contrast / saturation / hue jitter of strength ``s`` in random order, applied when
    the draw ``color_jitter`` exceeds ``p`` (dacs_transforms.py:42-59)."""
    if data is None or data.shape[1] != 3 or not color_jitter > p:
        return data, target
    x = denorm(data)
    order = [0, 1, 2, 3]
    random.shuffle(order)
    for t in order:
        f = random.uniform(1 - s, 1 + s)
        if t == 0:
            x = (x * f).clamp(0, 1)
        elif t == 1:
            m = _gray(x).mean(dim=(1, 2, 3), keepdim=True)
            x = ((x - m) * f + m).clamp(0, 1)
        elif t == 2:
            g = _gray(x)
            x = ((x - g) * f + g).clamp(0, 1)
        else:
            h = random.uniform(-s, s) * 2 * math.pi if s <= 0.5 else 0.0
            c, sn = math.cos(h), math.sin(h)
            # rotate the chroma plane of YIQ by h
            t_yiq = x.new_tensor([[0.299, 0.587, 0.114], [0.596, -0.274, -0.322], [0.211, -0.523, 0.312]])
            rot = x.new_tensor([[1, 0, 0], [0, c, -sn], [0, sn, c]])
            m = torch.linalg.inv(t_yiq) @ rot @ t_yiq
            x = torch.einsum('ij,bjhw->bihw', m, x).clamp(0, 1)
    return renorm(x), target


def gaussian_blur(blur, data=None, target=None):
    """Separable gaussian blur, sigma ~ U(0.15, 1.15), kernel ~ 10 % of the image side
    (dacs_transforms.py:62-78), reflect border."""
    if data is None or data.shape[1] != 3 or not blur > 0.5:
        return data, target
    sigma = random.uniform(0.15, 1.15)

    def ksize(n):
        c = math.ceil(0.1 * n)
        return int(math.floor(c - 0.5 + c % 2))

    def kernel1d(k):
        ax = torch.arange(k, device=data.device, dtype=torch.float32) - (k - 1) / 2.0
        w = torch.exp(-(ax ** 2) / (2 * sigma * sigma))
        return (w / w.sum()).to(data.dtype)

    ky, kx = ksize(data.shape[2]), ksize(data.shape[3])
    # the gaussian is < 1e-12 beyond 8 sigma: truncate the (up to 103-tap) kernel there
    ky, kx = min(ky, 2 * int(8 * sigma + 1) + 1), min(kx, 2 * int(8 * sigma + 1) + 1)
    C = data.shape[1]
    wy = kernel1d(ky).view(1, 1, ky, 1).repeat(C, 1, 1, 1)
    wx = kernel1d(kx).view(1, 1, 1, kx).repeat(C, 1, 1, 1)
    x = F.pad(data, (kx // 2, kx // 2, ky // 2, ky // 2), mode='reflect')
    x = F.conv2d(F.conv2d(x, wy, groups=C), wx, groups=C)
    return x, target


def strong_transform(param, data=None, target=None):
    assert data is not None or target is not None
    data, target = one_mix(mask=param['mix'], data=data, target=target)
    data, target = color_jitter(color_jitter=param['color_jitter'], s=param['color_jitter_s'],
                                p=param['color_jitter_p'], data=data, target=target)
    data, target = gaussian_blur(blur=param['blur'], data=data, target=target)
    return data, target
