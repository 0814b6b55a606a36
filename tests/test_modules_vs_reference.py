"""Host-side modules of refign_b200 vs the REAL reference modules (imported from /root/reference
through tests/refshim.py) with shared ``state_dict``s, on CPU.  The operator layer is routed to the
CPU oracle (tests/cpu_ops.py), so what is checked here is the host logic: state_dict compatibility
(strict load both ways), module wiring, channels-last token pipeline, BN folding, scale bookkeeping
in the UAWarpC head, refine/align glue.  Build container only (``needs_reference``).

Tolerance: 1e-3 relative (north_star) -- in practice 1e-5, everything is fp32.
"""
import types

import pytest
import torch

import refshim
from cpu_ops import cpu_ops


@pytest.fixture(autouse=True)
def _cpu_operator_layer():
    """Every test of this module runs the host-side modules on CPU: route refign_b200.ops to the oracle."""
    with cpu_ops():
        yield

pytestmark = pytest.mark.needs_reference


def close(a, b, rtol=1e-3, atol=2e-5):
    a, b = a.detach().float(), b.detach().float()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs()
    assert bool((err <= atol + rtol * b.abs()).all()), "max err %.3e, max ref %.3e" % (err.max(), b.abs().max())


@pytest.fixture(scope="module")
def ref():
    refshim.install()
    import models.backbones as rb
    import models.heads as rh
    import models.segmentation_model as rs
    import models.alignment_model as ra
    return types.SimpleNamespace(b=rb, h=rh, s=rs, a=ra)


def _randomise_bn(m, seed=0):
    g = torch.Generator().manual_seed(seed)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
            mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
            mod.bias.data.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)


def test_mit_matches_reference(ref):
    from refign_b200 import MixVisionTransformer
    torch.manual_seed(0)
    r = ref.b.MixVisionTransformer('mit_b0', drop_path_rate=0.0).eval()
    m = MixVisionTransformer('mit_b0', drop_path_rate=0.0).eval()
    assert list(m.state_dict().keys()) == list(r.state_dict().keys())
    m.load_state_dict(r.state_dict(), strict=True)
    r.load_state_dict(m.state_dict(), strict=True)
    x = torch.randn(2, 3, 64, 96)
    with torch.no_grad():
        for a, b in zip(m(x), r(x)):
            close(a, b)
    # gradients through the token pipeline
    x.requires_grad_(True)
    la = sum(o.square().mean() for o in m(x))
    ga = torch.autograd.grad(la, [x] + list(m.parameters()))
    lb = sum(o.square().mean() for o in r(x))
    gb = torch.autograd.grad(lb, [x] + list(r.parameters()))
    for a, b in zip(ga, gb):
        close(a, b, atol=1e-6)


def test_daformer_head_matches_reference(ref):
    from refign_b200 import DAFormerHead
    torch.manual_seed(1)
    args = dict(in_channels=[32, 64, 160, 256], in_index=[0, 1, 2, 3], num_classes=19,
                input_transform='multiple_select', dropout_ratio=0.0)
    r = ref.h.DAFormerHead(**args)
    m = DAFormerHead(**args)
    assert list(m.state_dict().keys()) == list(r.state_dict().keys())
    _randomise_bn(r)
    m.load_state_dict(r.state_dict(), strict=True)
    feats = [torch.randn(2, c, 32 // s, 48 // s) for c, s in zip(args['in_channels'], (1, 2, 4, 8))]
    for mode in (False, True):  # eval (running stats) and train (batch stats, as the student/teacher run)
        r.train(mode)
        m.train(mode)
        with torch.no_grad():
            close(m(feats), r(feats))


def test_vgg_matches_reference(ref):
    from refign_b200 import VGG
    torch.manual_seed(2)
    r = ref.b.VGG('vgg16', out_indices=[2, 3, 4]).eval()
    m = VGG('vgg16', out_indices=[2, 3, 4]).eval()
    assert m.layer_indices == r.layer_indices == [10, 17, 24]
    m.load_state_dict(r.state_dict(), strict=True)
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        for idx in ([-3, -2], [-2, -1], None):
            for a, b in zip(m(x, extract_only_indices=idx), r(x, extract_only_indices=idx)):
                close(a, b)


def _alignment_pair(ref, seed=3):
    from refign_b200 import VGG, UAWarpCHead
    torch.manual_seed(seed)
    rv = ref.b.VGG('vgg16', out_indices=[2, 3, 4]).eval()
    rh = ref.h.UAWarpCHead(in_index=[0, 1], input_transform='multiple_select', estimate_uncertainty=True).eval()
    _randomise_bn(rh)
    mv = VGG('vgg16', out_indices=[2, 3, 4]).eval()
    mh = UAWarpCHead(in_index=[0, 1], input_transform='multiple_select', estimate_uncertainty=True).eval()
    assert sorted(mh.state_dict().keys()) == sorted(rh.state_dict().keys())
    mv.load_state_dict(rv.state_dict(), strict=True)
    mh.load_state_dict(rh.state_dict(), strict=True)
    return rv, rh, mv, mh


def test_alignment_model_forward_matches_reference(ref):
    from refign_b200 import AlignmentModel
    rv, rh, mv, mh = _alignment_pair(ref)
    rm = ref.a.AlignmentModel.__new__(ref.a.AlignmentModel)
    torch.nn.Module.__init__(rm)
    rm.alignment_backbone, rm.alignment_head = rv, rh
    mm = AlignmentModel(alignment_backbone=mv, alignment_head=mh).eval()
    torch.manual_seed(30)
    img_i = torch.randn(1, 3, 256, 320)
    img_j = img_i.roll((3, -5), (2, 3)) + 0.05 * torch.randn(1, 3, 256, 320)
    with torch.no_grad(), cpu_ops():
        f_m, u_m = mm(img_i, img_j)
        f_r, u_r = rm.forward(img_i, img_j)
    close(f_m, f_r, atol=1e-3)   # flow in pixels: 1e-3 px absolute + 1e-3 relative
    close(u_m, u_r, atol=1e-4)


def test_align_refine_match_reference(ref):
    from refign_b200 import DomainAdaptationSegmentationModel as M
    rv, rh, mv, mh = _alignment_pair(ref, seed=4)
    torch.manual_seed(40)
    img_t = torch.randn(1, 3, 256, 256)
    img_r = img_t.roll((-4, 6), (2, 3)) + 0.05 * torch.randn(1, 3, 256, 256)
    lt, lr = torch.randn(1, 19, 256, 256) * 3, torch.randn(1, 19, 256, 256) * 3
    R = ref.s.DomainAdaptationSegmentationModel
    rself = types.SimpleNamespace(alignment_backbone=rv, alignment_head=rh, gamma=0.25, disable_M=False,
                                  disable_P=False, eta=R.eta)
    mself = types.SimpleNamespace(alignment_backbone=mv, alignment_head=mh, gamma=0.25, disable_M=False,
                                  disable_P=False, eta=M.eta, _fused_pseudo=None)
    with torch.no_grad(), cpu_ops():
        w_r, m_r, c_r = R.align(rself, lr, img_r, img_t)
        p_r = R.refine(rself, lt, w_r, m_r, c_r)
        w_m, m_m, c_m = M.align(mself, lr, img_r, img_t)
        # refine is discontinuous in its inputs (argmax -> M), so it is compared on IDENTICAL inputs
        p_m = M.refine(mself, lt, w_r, m_r, c_r)
        # fused variant used inside training_step: log-variance handed to the refine kernel
        w2, m2, lv = M.align(mself, lr, img_r, img_t, return_logvar=True)
        p2 = M.refine(mself, lt, w2, m2, None, logvar=lv)
        fused = mself._fused_pseudo
        p2_ref = M.refine(mself, lt, w_m, m_m, c_m)
    assert (m_m != m_r).float().mean() < 1e-3   # mask flips only where x+flow sits on the border to 1e-3 px
    close(c_m, c_r, atol=1e-4)
    same = (m_m == m_r).unsqueeze(1)
    close(torch.where(same, w_m, w_r), w_r, atol=5e-3)
    close(p_m, p_r, atol=1e-6)
    close(p2, p2_ref, atol=1e-6)
    label = fused[1]
    assert torch.equal(label, p2.argmax(1))
    close(M.eta(lt), R.eta(lt), atol=1e-6)
