"""Fused per-pixel patch CNN of UncertaintyModule (csrc/uncertainty_cnn.cu through ops.uncertainty_patch_cnn) against
the module's own layer-by-layer fp32 path (reference models/modules.py:534-561): search sizes 9 and 16 (with the 2x2
max-pool), pixel counts that are not multiples of the 16-pixel CTA group, BatchNorm statistics away from identity.
Tolerance: bf16 activations between the four layers, fp32 accumulation -> 2e-2 of the output scale (the library path
under bf16 autocast has the same error against fp32)."""
import pytest
import torch

import refign_b200 as P
from refign_b200 import ops
from refign_b200.modules import UncertaintyModule

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _module(search, feed):
    torch.manual_seed(search + feed)
    m = UncertaintyModule(in_channels=1, search_size=search, feed_in_previous=bool(feed)).to(DEV).eval()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0.0, 0.2)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.7, 1.3)
                mod.bias.normal_(0.0, 0.2)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


@pytest.mark.parametrize("spec", [(9, 2, 32, 32), (9, 1, 13, 11), (16, 2, 16, 16), (16, 1, 5, 7), (9, 2, 128, 128)])
def test_patch_cnn_matches_layerwise_fp32(spec):
    s, B, H, W = spec
    m = _module(s, 0)
    corr = torch.rand(B, s * s, H, W, device=DEV)
    with torch.no_grad():
        x = corr.permute(0, 2, 3, 1).reshape(B * H * W, 1, s, s)
        x = m.conv_0(x)
        if s == 16:
            x = m.maxpool(x)
        want = m.predict_uncertainty(m.conv_2(m.conv_1(x))).reshape(B, H, W, 6).permute(0, 3, 1, 2)
        got = ops.uncertainty_patch_cnn(corr, m._fused_params(), s, 0.1)
    assert got.shape == want.shape and got.dtype == torch.bfloat16
    scale = float(want.abs().max())
    err = float((got.float() - want).abs().max())
    assert err <= 2e-2 * scale, (err, scale)


@pytest.mark.parametrize("search", [9, 16])
def test_uncertainty_module_forward_takes_the_fused_kernel(search):
    feed = search == 9
    m = _module(search, int(feed))
    B, H, W = 2, 24, 40
    corr = torch.rand(B, search * search, H, W, device=DEV)
    feat = torch.randn(B, 32, H, W, device=DEV)
    extra = (torch.randn(B, 1, H, W, device=DEV), torch.randn(B, 2, H, W, device=DEV)) if feed else ()
    with torch.no_grad():
        want = m(corr, feat, *extra)                       # fp32, layer by layer (library convolutions)
        timer = ops.KernelTimer()
        ops.set_timer(timer)
        try:
            with torch.autocast('cuda', dtype=torch.bfloat16):
                got = m(corr, feat, *extra)
        finally:
            ops.set_timer(None)
    assert "uncertainty_cnn" in {r[0] for r in timer.records}
    scale = float(want.abs().max())
    assert float((got.float() - want).abs().max()) <= 3e-2 * scale
