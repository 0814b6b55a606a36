"""Host mirror of the DACS strong transform (refign_b200/dacs_transforms.py, the CPU path of get_dacs_mix) against
the restatement of kornia 0.5.8 in oracle/kornia_058.py: same explicit parameters -> same image, label and weight.
(kornia itself is absent: parity unpinned against a live kornia, see oracle/kornia_058.py.)"""
import math
import random

import torch

from oracle import kornia_058 as K
from refign_b200 import dacs_transforms as D


def _oracle_args(row, H, W):
    jitter = blur = None
    if float(row[0]):
        jitter = ([int(v) for v in row[1:5]], float(row[5]) + 1.0, float(row[6]), float(row[7]), float(row[8]) / (2 * math.pi))
    return jitter, blur


def test_color_jitter_mirror_matches_kornia_restatement():
    random.seed(1)
    torch.manual_seed(1)
    for trial in range(6):
        P = D.draw_strong_params(1, 48, 40, 0.9, 0.25 if trial % 2 else {'brightness': 0.4, 'contrast': 0.3, 'saturation': 0.5, 'hue': 0.1},
                                 0.2, 0.0)
        x = torch.randn(1, 3, 48, 40)
        jitter, _ = _oracle_args(P[0], 48, 40)
        want = (K.color_jitter(x * K._STD + K._MEAN, *jitter) - K._MEAN) / K._STD
        got = D.apply_color_jitter(x, P[0])
        assert torch.allclose(got, want, rtol=0, atol=1e-6), float((got - want).abs().max())


def test_blur_mirror_matches_full_size_kornia_kernel():
    """The truncated half-kernel form equals kornia's full (~10 % of the side) 2-D gaussian with reflect border."""
    torch.manual_seed(2)
    for sigma, (H, W) in ((0.15, (40, 56)), (0.7, (96, 64)), (1.15, (128, 128))):
        ky, kx = D.blur_kernel_size(H), D.blur_kernel_size(W)
        assert ky % 2 == 1 and kx % 2 == 1
        row = torch.zeros(D.PARAM_STRIDE)
        ry, wy = D._half_kernel(ky, sigma)
        rx, wx = D._half_kernel(kx, sigma)
        row[9], row[10], row[11] = 1.0, ry, rx
        row[12:12 + ry + 1] = torch.tensor(wy)
        row[29:29 + rx + 1] = torch.tensor(wx)
        x = torch.randn(1, 3, H, W)
        want = K.gaussian_blur2d(x, (ky, kx), sigma)
        got = D.apply_gaussian_blur(x, row)
        assert got.shape == want.shape and torch.allclose(got, want, rtol=0, atol=2e-6), float((got - want).abs().max())


def test_draw_strong_params_distributions():
    random.seed(3)
    P = D.draw_strong_params(64, 512, 512, 0.9, 0.25, 0.2, 0.9)
    assert bool((P[:, 0] == 1).all()) and bool((P[:, 9] == 1).all())
    assert sorted(int(v) for v in P[0, 1:5]) == [0, 1, 2, 3]
    assert float(P[:, 5].abs().max()) <= 0.25 + 1e-6 and 0.75 - 1e-6 <= float(P[:, 6].min()) and float(P[:, 6].max()) <= 1.25 + 1e-6
    assert float(P[:, 8].abs().max()) <= 0.25 * 2 * math.pi + 1e-6
    sums = P[:, 12] + 2 * P[:, 13:29].sum(1)
    assert torch.allclose(sums, torch.ones(64), atol=1e-6)
    off = D.draw_strong_params(2, 64, 64, 0.1, 0.25, 0.2, 0.3)      # draws below the thresholds: everything off
    assert float(off.abs().sum()) == 0.0
