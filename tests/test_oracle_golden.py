"""CPU: the oracle (oracle/refign_oracle.c) against the golden vectors produced by
the reference's own code (tests/golden/make_golden.py).  This is what pins the
oracle; the GPU tests then compare the CUDA kernels with the oracle."""
import numpy as np
import torch

import oracle


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_local_corr_fwd_bwd_bit_exact(golden):
    g = golden("ops_local_corr")
    for i in range(int(g["ncases"])):
        B, C, H, W, k, P, s, pad, dil, dp = [int(v) for v in g[f"c{i}_spec"]]
        a, b = T(g[f"c{i}_in1"]), T(g[f"c{i}_in2"])
        out = oracle.local_corr_fwd(a, b, k, P, s, pad, dil, dp)
        assert torch.equal(out, T(g[f"c{i}_out"])), f"case {i}: forward differs from reference"
        ga, gb = oracle.local_corr_bwd(a, b, T(g[f"c{i}_gout"]), k, P, s, pad, dil, dp)
        assert torch.equal(ga, T(g[f"c{i}_gin1"])) and torch.equal(gb, T(g[f"c{i}_gin2"])), f"case {i}: backward"


def test_corr_layers(golden):
    g = golden("ops_corr_layers")
    loc = oracle.local_corr_layer(T(g["src"]), T(g["trg"]), 9)
    assert torch.allclose(loc, T(g["local"]), rtol=0, atol=1e-6)
    glob = oracle.global_corr(T(g["gsrc"]), T(g["gtrg"]), mutual=True)
    assert torch.allclose(glob, T(g["glob"]), rtol=1e-4, atol=1e-6)
    glob = oracle.global_corr(T(g["gsrc"]), T(g["gtrg"]), mutual=False)
    assert torch.allclose(glob, T(g["glob_nomm"]), rtol=1e-4, atol=1e-6)


def test_warp_mask_cert(golden):
    g = golden("ops_warp")
    x = T(g["x"])
    out, mask = oracle.warp(x, T(g["flow"]), return_mask=True)
    assert torch.equal(mask, T(g["mask"]))  # boolean output: bit-exact
    assert torch.allclose(out, T(g["out"]), rtol=1e-5, atol=1e-5)
    out, mask = oracle.warp(x[:1], T(g["flow_big"]), return_mask=True)
    assert torch.equal(mask, T(g["mask_big"]))
    assert torch.allclose(out, T(g["out_big"]), rtol=1e-5, atol=1e-5)
    out, mask = oracle.warp(x, torch.zeros_like(T(g["flow"])), return_mask=True)
    assert torch.equal(out, T(g["out_zero"])) and bool(mask.all())
    assert torch.allclose(oracle.cert(T(g["logvar"])), T(g["cert"]), rtol=0, atol=3e-7)


def test_refine_labels_bit_exact(golden):
    g = golden("ops_refine")
    lt, lr, m, ce = T(g["lt"]), T(g["lr"]), T(g["mask"]), T(g["certs"])
    cfgs = {"full": (False, False, m, ce), "noM": (True, False, m, ce), "noP": (False, True, m, ce),
            "bare": (False, False, None, None)}
    for tag, (dM, dP, mask, certs) in cfgs.items():
        probs, label, maxp, trust = oracle.refine(lt, lr, mask, certs=certs, gamma=0.25, disable_M=dM,
                                                  disable_P=dP)
        assert torch.equal(label, T(g[f"{tag}_label"])), tag          # int64 pseudo-label: bit-exact
        assert torch.allclose(probs, T(g[f"{tag}_probs"]), rtol=0, atol=3e-7), tag
        assert torch.allclose(maxp, T(g[f"{tag}_maxprob"]), rtol=0, atol=3e-7), tag
    assert torch.allclose(trust, T(g["trust"]), rtol=1e-6, atol=0)


def test_exact_transcendentals():
    x = torch.linspace(-100, 88, 200001)
    ref = torch.exp(x.double())
    got = oracle.exact_expf(x).double()
    ok = ref > 1e-37
    assert ((got - ref).abs()[ok] / ref[ok]).max() < 3e-7
    x = torch.linspace(0.5, 40, 100001)
    assert (oracle.exact_logf(x).double() - torch.log(x.double())).abs().max() < 5e-7


def test_upsample_ce_loss_tail(golden):
    """oracle.upsample_ce (C restatement of F.interpolate + PixelWeightedCrossEntropyLoss, gradient included) against
    the loss values and logit gradients the reference's own loss class produced (tests/golden/make_golden_loss.py)."""
    g = golden("ops_upsample_ce")
    for i in range(int(g["ncases"])):
        w = T(g[f"c{i}_weight"])
        got, grad = oracle.upsample_ce(T(g[f"c{i}_logits"]), T(g[f"c{i}_target"]), w if w.numel() else None, 255,
                                       return_grad=True)
        want = float(g[f"c{i}_loss"])
        assert abs(float(got) - want) <= 2e-6 * max(1.0, abs(want)), (i, float(got), want)
        g_want = T(g[f"c{i}_grad"])
        assert torch.allclose(grad, g_want, rtol=1e-4, atol=2e-6 * float(g_want.abs().max())), i
