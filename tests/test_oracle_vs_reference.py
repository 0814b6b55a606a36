"""CPU, build container only: the oracle against the LIVE reference on fresh random
inputs (the golden fixtures are a frozen subset of this)."""
import types

import pytest
import torch

import oracle
import refshim

pytestmark = pytest.mark.needs_reference


@pytest.fixture(scope="module")
def ref():
    refshim.install()
    import helpers.matching_utils as mu
    import models.modules as mm
    from models.segmentation_model import DomainAdaptationSegmentationModel as M
    return types.SimpleNamespace(mu=mu, mm=mm, M=M, ext=refshim.ref_ext())


@pytest.mark.parametrize("spec", [(2, 32, 20, 24, 1, 9, 1, 0, 1, 1), (1, 6, 11, 9, 3, 3, 2, 1, 2, 3),
                                  (1, 4, 7, 7, 1, 6, 1, 0, 1, 1)])
def test_local_corr(ref, spec):
    B, C, H, W, k, P, s, pad, dil, dp = spec
    torch.manual_seed(sum(spec))
    a, b = torch.randn(B, C, H, W), torch.randn(B, C, H, W)
    r = ref.ext.forward(a, b, k, k, P, P, pad, pad, dil, dil, dp, dp, s, s)
    assert torch.equal(r, oracle.local_corr_fwd(a, b, k, P, s, pad, dil, dp))
    g = torch.randn_like(r)
    r1, r2 = ref.ext.backward(a, b, g, k, k, P, P, pad, pad, dil, dil, dp, dp, s, s)
    o1, o2 = oracle.local_corr_bwd(a, b, g, k, P, s, pad, dil, dp)
    assert torch.equal(r1, o1) and torch.equal(r2, o2)


def test_refine_512(ref):
    torch.manual_seed(7)
    lt, lr = torch.randn(2, 19, 256, 256) * 3, torch.randn(2, 19, 256, 256) * 3
    mask, ce = torch.rand(2, 256, 256) > 0.15, torch.rand(2, 1, 256, 256)
    self = types.SimpleNamespace(gamma=0.25, disable_M=False, disable_P=False, eta=ref.M.eta)
    r = ref.M.refine(self, lt, lr, mask, ce)
    rp, rl = torch.max(r, dim=1)
    probs, label, maxp, _ = oracle.refine(lt, lr, mask, certs=ce)
    assert torch.equal(rl, label)
    assert torch.allclose(r, probs, rtol=0, atol=3e-7)


def test_warp_and_layers(ref):
    torch.manual_seed(3)
    x, flo = torch.randn(1, 8, 33, 47), torch.randn(1, 2, 33, 47) * 5
    r, rm = ref.mu.warp(x, flo, return_mask=True)
    o, om = oracle.warp(x, flo, return_mask=True)
    assert torch.equal(rm, om) and torch.allclose(r, o, atol=1e-5)
    s = torch.nn.functional.normalize(torch.randn(1, 48, 16, 16), dim=1)
    t = torch.nn.functional.normalize(torch.randn(1, 48, 16, 16), dim=1)
    assert torch.allclose(ref.mm.GlobalFeatureCorrelationLayer()(s, t), oracle.global_corr(s, t), rtol=1e-4, atol=1e-6)
    assert torch.allclose(ref.mm.LocalFeatureCorrelationLayer(9)(s, t), oracle.local_corr_layer(s, t), atol=1e-6)
