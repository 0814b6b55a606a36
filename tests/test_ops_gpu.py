"""GPU parity tests: the sm_100a kernels, called through the C-ABI (refign_b200.ops ->
ctypes -> librefign_b200.so), against the CPU oracle on the same seeded inputs, against
the committed golden vectors of the reference, and through size-independent properties
at BASELINE sizes.

Tolerances: integer / boolean outputs (pseudo-label, warp mask) bit-exact; floating
point within 1e-3 relative (north_star), in practice ~1e-6 since everything is fp32.
"""
import numpy as np
import pytest
import torch

import oracle
from refign_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-3  # BASELINE.json north_star: "within 1e-3 rel fp for logits/flow"


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(got, want, atol=1e-5, rtol=RTOL):
    got, want = got.detach().cpu().float(), want.detach().cpu().float()
    assert got.shape == want.shape, (got.shape, want.shape)
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    assert bool((err <= tol).all()), "max abs err %.3e (max |want| %.3e)" % (err.max().item(), want.abs().max().item())


def unit(x):
    return torch.nn.functional.normalize(x, p=2, dim=1)


# ----------------------------------------------------------------------------- local corr
LC_SPECS = [  # B C H W k P stride pad dil dil_patch
    (2, 16, 12, 16, 1, 9, 1, 0, 1, 1),      # tiled kernel, partial tiles
    (1, 128, 64, 64, 1, 9, 1, 0, 1, 1),     # BASELINE config 1 shape
    (2, 37, 19, 44, 1, 9, 1, 0, 1, 1),      # C not a multiple of the pipeline chunk, ragged tile
    (1, 8, 10, 13, 1, 9, 1, 0, 1, 1),       # W % 4 != 0 -> generic kernel
    (1, 7, 16, 16, 1, 5, 1, 0, 1, 1),
    (1, 9, 8, 12, 1, 7, 1, 0, 1, 1),
    (1, 6, 9, 8, 1, 3, 1, 0, 1, 1),
    (1, 4, 8, 8, 1, 1, 1, 0, 1, 1),         # patch 1 == plain channel dot product
    (1, 4, 8, 8, 1, 4, 1, 0, 1, 1),         # even patch
    (1, 5, 9, 11, 3, 5, 2, 1, 1, 2),        # kernel 3, stride 2, pad 1, dilated patch
    (1, 3, 10, 8, 2, 3, 1, 2, 2, 1),
    (1, 1, 1, 4, 1, 9, 1, 0, 1, 1),         # single row
    (1, 16, 20, 70, 1, 19, 1, 0, 1, 1),     # wide-patch kernel (sweep stress point d = 9), ragged 64-wide tiles
    (2, 9, 12, 40, 1, 33, 1, 0, 1, 1),      # wide-patch kernel, d = 16, C not a multiple of the chunk
    (1, 8, 9, 13, 1, 11, 1, 0, 1, 1),       # wide-patch kernel, odd W (scalar stores)
]


@pytest.mark.parametrize("spec", LC_SPECS)
def test_local_corr_fwd_bwd_vs_oracle(spec):
    B, C, H, W, k, P, s, pad, dil, dp = spec
    torch.manual_seed(100 + sum(spec))
    a, b = torch.randn(B, C, H, W), torch.randn(B, C, H, W)
    want = oracle.local_corr_fwd(a, b, k, P, s, pad, dil, dp)
    ad, bd = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    got = ops.spatial_correlation_sample(ad, bd, k, P, s, pad, dil, dp)
    close(got, want, atol=1e-4 * max(1.0, C ** 0.5))
    g = torch.randn_like(want)
    got.backward(g.to(DEV))
    wa, wb = oracle.local_corr_bwd(a, b, g, k, P, s, pad, dil, dp)
    close(ad.grad, wa, atol=1e-4 * P)
    close(bd.grad, wb, atol=1e-4 * P)


@pytest.mark.parametrize("P", [19, 33])
def test_local_corr_wide_patch_layer_vs_oracle(P):
    """LocalFeatureCorrelationLayer semantics (ReLU + L2-norm over the P*P displacements) at the sweep's wide
    patches: the wide-patch kernel + the generic ReLU / L2-norm pass vs the oracle."""
    torch.manual_seed(P)
    s, t = unit(torch.randn(2, 32, 24, 72)), unit(torch.randn(2, 32, 24, 72))
    want = oracle.local_corr_layer(s, t, P)
    got = ops.local_correlation_relu_l2norm(s.to(DEV), t.to(DEV), P)
    close(got, want, atol=2e-6)
    n = got.norm(dim=1)
    assert bool(((n - 1).abs() < 1e-4).logical_or(n == 0).all())


def test_local_corr_golden(golden):
    g = golden("ops_local_corr")
    for i in range(int(g["ncases"])):
        B, C, H, W, k, P, s, pad, dil, dp = [int(v) for v in g[f"c{i}_spec"]]
        a, b = T(g[f"c{i}_in1"]).to(DEV).requires_grad_(True), T(g[f"c{i}_in2"]).to(DEV).requires_grad_(True)
        out = ops.spatial_correlation_sample(a, b, k, P, s, pad, dil, dp)
        close(out, T(g[f"c{i}_out"]), atol=2e-5 * C)
        out.backward(T(g[f"c{i}_gout"]).to(DEV))
        close(a.grad, T(g[f"c{i}_gin1"]), atol=1e-4 * P)
        close(b.grad, T(g[f"c{i}_gin2"]), atol=1e-4 * P)


def test_local_corr_layer_fused_and_grad(golden):
    g = golden("ops_corr_layers")
    src, trg = T(g["src"]).to(DEV), T(g["trg"]).to(DEV)
    close(ops.local_correlation_relu_l2norm(src, trg, 9), T(g["local"]), atol=2e-6)
    # fused epilogue == unfused op + torch relu/normalize, values and gradients
    torch.manual_seed(5)
    s = unit(torch.randn(2, 24, 16, 20)).to(DEV).requires_grad_(True)
    t = unit(torch.randn(2, 24, 16, 20)).to(DEV).requires_grad_(True)
    y = ops.local_correlation_relu_l2norm(s, t, 9)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    gs, gt = s.grad.clone(), t.grad.clone()
    s.grad = t.grad = None
    c = ops.spatial_correlation_sample(t, s, patch_size=9).view(2, 81, 16, 20)
    y2 = torch.nn.functional.normalize(torch.relu(c), p=2, dim=1)
    (y2 * w).sum().backward()
    close(y, y2, atol=2e-6)
    close(gs, s.grad, atol=1e-5)
    close(gt, t.grad, atol=1e-5)
    # and against the CPU autograd of the oracle formula (zero-pad-and-shift restatement)
    sc, tc = s.detach().cpu().requires_grad_(True), t.detach().cpu().requires_grad_(True)
    pad = torch.nn.functional.pad(sc, (4, 4, 4, 4))
    planes = [(tc * pad[:, :, ph:ph + 16, pw:pw + 20]).sum(1) for ph in range(9) for pw in range(9)]
    y3 = torch.nn.functional.normalize(torch.relu(torch.stack(planes, 1)), p=2, dim=1)
    (y3 * w.cpu()).sum().backward()
    close(y, y3, atol=2e-6)
    close(gs, sc.grad, atol=1e-5)
    close(gt, tc.grad, atol=1e-5)


def test_local_corr_properties_full_size():
    """BASELINE sizes (level-1 volume of a 1024^2 pair: B2 C128 256x256 P9): swap symmetry,
    linearity and agreement between the tiled and the generic kernel on a crop."""
    torch.manual_seed(11)
    B, C, H, W, P = 2, 128, 256, 256, 9
    a, b = unit(torch.randn(B, C, H, W, device=DEV)), unit(torch.randn(B, C, H, W, device=DEV))
    ab = ops.spatial_correlation_sample(a, b, patch_size=P)
    ba = ops.spatial_correlation_sample(b, a, patch_size=P)
    # corr(a,b)[ph,pw,y,x] == corr(b,a)[8-ph,8-pw,y+dy,x+dx]
    for ph, pw in [(0, 0), (4, 4), (2, 7), (8, 3)]:
        dy, dx = ph - 4, pw - 4
        ys, xs = slice(max(0, -dy), H - max(0, dy)), slice(max(0, -dx), W - max(0, dx))
        ys2, xs2 = slice(max(0, dy), H - max(0, -dy)), slice(max(0, dx), W - max(0, -dx))
        close(ab[:, ph, pw, ys, xs], ba[:, 8 - ph, 8 - pw, ys2, xs2], atol=1e-6)
    # centre displacement is the plain dot product; unit-norm features => |corr| <= 1
    # (this shape runs the tensor-core banded GEMM: bf16 hi/lo split operands, three MMAs -> <= 1e-5 absolute on
    #  unit-norm features, 100x inside north_star's 1e-3; the exact-fp32 FFMA tiles hold 2e-6, see the A/B test below)
    close(ab[:, 4, 4], (a * b).sum(1), atol=1e-5)
    assert ab.abs().max().item() <= 1.0 + 1e-5
    # out-of-image displacements are exactly zero
    assert ab[:, 0, :, :4, :].abs().max().item() == 0.0 and ab[:, :, 8, :, -4:].abs().max().item() == 0.0
    # linearity in the second argument
    b2 = torch.randn_like(b)
    lin = ops.spatial_correlation_sample(a, 0.5 * b + 2.0 * b2, patch_size=P)
    # (b2 is not normalised: |values| ~ 2 sqrt(C); the split-operand error scales with the operands, 2e-5 of that scale)
    close(lin, 0.5 * ab + 2.0 * ops.spatial_correlation_sample(a, b2, patch_size=P), atol=2e-5 * max(1.0, float(lin.abs().max())))
    # a crop (with halo) through the oracle
    want = oracle.local_corr_fwd(a[:1, :, 100:132, 60:100].cpu(), b[:1, :, 100:132, 60:100].cpu(), patch_size=P)
    close(ab[:1, :, :, 104:128, 64:96], want[:, :, :, 4:28, 4:36], atol=1e-5)


@pytest.mark.parametrize("shape", [(2, 128, 256, 256), (1, 128, 64, 64), (2, 64, 100, 68), (1, 48, 65, 132)])
@pytest.mark.parametrize("fused", [False, True])
def test_local_corr_tensor_core_vs_exact_tiles(shape, fused, monkeypatch):
    """The tcgen05 banded-GEMM kernel (csrc/local_corr_tc.cu, default for the 9 x 9 volume on maps >= 64 x 64) against
    the exact-fp32 FFMA tiles (RF_LOCAL_CORR_TC=0) on unit-norm features: ragged tiles, image borders (TMA zero fill =
    the correlation's padding), fused ReLU + L2-norm.  Written bound: 1e-5 absolute (values in [-1, 1])."""
    B, C, H, W = shape
    torch.manual_seed(B * C + H + W)
    a, b = unit(torch.randn(B, C, H, W, device=DEV)), unit(torch.randn(B, C, H, W, device=DEV))
    run = (lambda: ops.local_correlation_relu_l2norm(a, b, 9)) if fused else (lambda: ops.spatial_correlation_sample(a, b, patch_size=9))
    monkeypatch.setenv("RF_LOCAL_CORR_TC", "0")
    exact = run()
    monkeypatch.setenv("RF_LOCAL_CORR_TC", "1")
    timer = ops.KernelTimer()
    ops.set_timer(timer)
    try:
        got = run()
    finally:
        ops.set_timer(None)
    assert got.shape == exact.shape
    err = float((got - exact).abs().max())
    assert err <= 1e-5, err
    assert not torch.equal(got, exact) or C < 16      # (the two kernels really are different code paths)
    # zero padding is exact: displacements that leave the image
    g5 = got.view(B, 9, 9, H, W)
    assert float(g5[:, 0, :, :4, :].abs().max()) == 0.0 and float(g5[:, :, 8, :, -4:].abs().max()) == 0.0


def test_local_corr_errors():
    a = torch.randn(1, 4, 8, 8, device=DEV)
    with pytest.raises(RuntimeError):
        ops.spatial_correlation_sample(a, a[:, :2], patch_size=9)
    with pytest.raises(RuntimeError):
        ops.spatial_correlation_sample(a, a, kernel_size=11, patch_size=3)  # empty output
    with pytest.raises(RuntimeError):
        ops.spatial_correlation_sample(a[0], a[0], patch_size=3)


# ----------------------------------------------------------------------------- global corr
@pytest.mark.parametrize("shape", [(2, 64, 16, 16, 16, 16), (1, 40, 9, 7, 5, 11), (2, 512, 16, 16, 16, 16),
                                   (1, 128, 64, 64, 64, 64)])
@pytest.mark.parametrize("mm", [True, False])
def test_global_corr_vs_oracle(shape, mm):
    B, C, Hs, Ws, Ht, Wt = shape
    torch.manual_seed(sum(shape))
    s, t = unit(torch.randn(B, C, Hs, Ws)), unit(torch.randn(B, C, Ht, Wt))
    want = oracle.global_corr(s, t, mutual=mm)
    got = ops.global_correlation(s.to(DEV), t.to(DEV), cyclic_consistency=mm, use_tensor_cores=0)
    close(got, want, atol=2e-6)
    raw = ops.global_correlation(s.to(DEV), t.to(DEV), cyclic_consistency=False, normalise=False, use_tensor_cores=0)
    close(raw, oracle.global_corr(s, t, mutual=False, normalise=False), atol=2e-6)


def test_global_corr_golden(golden):
    g = golden("ops_corr_layers")
    s, t = T(g["gsrc"]).to(DEV), T(g["gtrg"]).to(DEV)
    close(ops.global_correlation(s, t, use_tensor_cores=0), T(g["glob"]), atol=2e-6)
    close(ops.global_correlation(s, t, cyclic_consistency=False, use_tensor_cores=0), T(g["glob_nomm"]), atol=2e-6)


@pytest.mark.parametrize("shape", [(1, 128, 64, 64, 64, 64), (2, 256, 32, 32, 32, 32), (1, 128, 36, 28, 20, 52),
                                   (1, 32, 16, 12, 128, 128)])
@pytest.mark.parametrize("mm", [True, False])
def test_global_corr_tcgen05_vs_oracle(shape, mm):
    """tcgen05 kind::tf32 path: the operands are read as TF32 (10-bit mantissa) with fp32 accumulation,
    so a correlation of unit vectors carries ~2e-4 absolute error; mutual matching cubes the value
    (x3 relative error).  Tolerance: raw volume 1e-3 absolute; finished volume 2e-2 relative + 3e-3 of
    the largest entry.  (The model's own 16x16 case runs the exact-fp32 FFMA path; `-1` picks this
    path only for volumes of >= 2^20 entries.)"""
    B, C, Hs, Ws, Ht, Wt = shape
    torch.manual_seed(sum(shape) + 1)
    s, t = unit(torch.randn(B, C, Hs, Ws)), unit(torch.randn(B, C, Ht, Wt))
    raw = ops.global_correlation(s.to(DEV), t.to(DEV), cyclic_consistency=False, normalise=False, use_tensor_cores=1)
    close(raw, oracle.global_corr(s, t, mutual=False, normalise=False), atol=1e-3, rtol=0)
    want = oracle.global_corr(s, t, mutual=mm)
    got = ops.global_correlation(s.to(DEV), t.to(DEV), cyclic_consistency=mm, use_tensor_cores=1)
    close(got, want, atol=3e-3 * float(want.abs().max()), rtol=2e-2)
    # size-independent property: every target column of the finished volume has unit (or zero) norm
    n = got.norm(dim=1)
    assert bool(((n - 1).abs() < 1e-4).logical_or(n == 0).all())


@pytest.mark.parametrize("shape", [(1, 128, 64, 64, 64, 64), (2, 64, 32, 32, 32, 32), (1, 128, 36, 28, 20, 52),
                                   (1, 32, 16, 12, 128, 128), (3, 96, 40, 40, 24, 24), (2, 128, 50, 44, 36, 60),
                                   (1, 128, 80, 80, 80, 80)])
@pytest.mark.parametrize("mm", [True, False])
def test_global_corr_persistent_vs_oracle(shape, mm):
    """Persistent warp-specialised tcgen05 kind::tf32 path (use_tensor_cores=2; the automatic choice for
    volumes >= 2^20 entries with C <= 128): same tolerances as the single-tile tcgen05 path above.  The shapes
    cover one tile per CTA, many tiles per CTA (TMEM accumulator / operand ring wrap-around, source tile
    replaced inside a CTA's range), ragged last tiles in both dimensions, B > 1 and 1..4 K blocks."""
    B, C, Hs, Ws, Ht, Wt = shape
    torch.manual_seed(sum(shape) + 2)
    s, t = unit(torch.randn(B, C, Hs, Ws)), unit(torch.randn(B, C, Ht, Wt))
    raw = ops.global_correlation(s.to(DEV), t.to(DEV), cyclic_consistency=False, normalise=False, use_tensor_cores=2)
    close(raw, oracle.global_corr(s, t, mutual=False, normalise=False), atol=1e-3, rtol=0)
    want = oracle.global_corr(s, t, mutual=mm)
    got = ops.global_correlation(s.to(DEV), t.to(DEV), cyclic_consistency=mm, use_tensor_cores=2)
    close(got, want, atol=3e-3 * float(want.abs().max()), rtol=2e-2)
    n = got.norm(dim=1)
    assert bool(((n - 1).abs() < 1e-4).logical_or(n == 0).all())
    # the two tensor-core kernels run the same TF32 products: they agree far below the oracle tolerance
    old = ops.global_correlation(s.to(DEV), t.to(DEV), cyclic_consistency=mm, use_tensor_cores=1)
    close(got, old, atol=1e-5 * float(old.abs().max()), rtol=1e-4)


def test_global_corr_persistent_rejects_wide_channels():
    s = unit(torch.randn(1, 256, 8, 8)).to(DEV)
    with pytest.raises(RuntimeError):
        ops.global_correlation(s, s, use_tensor_cores=2)   # C > 128: the source tile does not stay resident


# ----------------------------------------------------------------------------- warp
def test_warp_golden_and_oracle(golden):
    g = golden("ops_warp")
    x = T(g["x"]).to(DEV)
    out, mask = ops.warp(x, T(g["flow"]).to(DEV), return_mask=True)
    assert mask.dtype == torch.bool and torch.equal(mask.cpu(), T(g["mask"]))
    close(out, T(g["out"]), atol=1e-5)
    out, mask = ops.warp(x[:1], T(g["flow_big"]).to(DEV), return_mask=True)
    assert torch.equal(mask.cpu(), T(g["mask_big"]))
    close(out, T(g["out_big"]), atol=1e-5)
    out, mask = ops.warp(x, torch.zeros(2, 2, 24, 40, device=DEV), return_mask=True)  # early-exit semantics
    assert torch.equal(out.cpu(), T(g["out_zero"])) and bool(mask.all())
    close(ops.estimate_probability_of_confidence_interval_of_mixture_density(T(g["logvar"]).to(DEV)), T(g["cert"]),
          atol=3e-7)


@pytest.mark.parametrize("shape", [(2, 19, 128, 160), (2, 256, 32, 32), (1, 3, 1, 7), (1, 5, 9, 1), (3, 130, 17, 23)])
def test_warp_bit_exact_vs_oracle(shape):
    B, C, H, W = shape
    torch.manual_seed(sum(shape))
    x, flo = torch.randn(B, C, H, W), torch.randn(B, 2, H, W) * 3 + 0.7
    flo[0, :, 0, 0] = 0
    wo, wm = oracle.warp(x, flo, return_mask=True)
    go, gm = ops.warp(x.to(DEV), flo.to(DEV), return_mask=True)
    assert torch.equal(gm.cpu(), wm)
    assert torch.equal(go.cpu(), wo), "same rounded operation sequence => bit-identical"
    assert torch.equal(ops.warp(x.to(DEV), flo.to(DEV)).cpu(), wo)


@pytest.mark.parametrize("kind", ["smooth", "noisy", "wild", "outside", "zero"])
def test_warp_flow_families_bit_exact_vs_oracle(kind):
    """Larger ragged shape (all channels per thread, one channel chunk): smooth flow, per-pixel noise, wild
    flow, everything out of the image, and the zero-flow early exit."""
    B, C, H, W = 3, 7, 181, 203
    torch.manual_seed(len(kind))
    x = torch.randn(B, C, H, W)
    if kind == "smooth":
        flo = torch.nn.functional.interpolate(torch.randn(B, 2, 6, 7) * 6, size=(H, W), mode="bilinear") + 1.3
    elif kind == "noisy":
        flo = torch.randn(B, 2, H, W) * 4 - 2.0
    elif kind == "wild":
        flo = torch.randn(B, 2, H, W) * 60
    elif kind == "outside":
        flo = torch.full((B, 2, H, W), 500.0)
        flo[1] = -400.0
        flo[2, :, :90] = torch.randn(2, 90, W)       # one image half valid
    else:
        flo = torch.zeros(B, 2, H, W)
    wo, wm = oracle.warp(x, flo, return_mask=True)
    go, gm = ops.warp(x.to(DEV), flo.to(DEV), return_mask=True)
    assert torch.equal(gm.cpu(), wm)
    assert torch.equal(go.cpu(), wo)


def test_warp_backward_vs_grid_sample():
    torch.manual_seed(9)
    B, C, H, W = 2, 6, 20, 28
    x, flo = torch.randn(B, C, H, W), torch.randn(B, 2, H, W) * 2.5
    w = torch.randn(B, C, H, W)
    xc, fc = x.clone().requires_grad_(True), flo.clone().requires_grad_(True)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    grid = torch.stack((2 * (xx + fc[:, 0]) / (W - 1) - 1, 2 * (yy + fc[:, 1]) / (H - 1) - 1), -1)
    (torch.nn.functional.grid_sample(xc, grid, align_corners=True) * w).sum().backward()
    xd, fd = x.to(DEV).requires_grad_(True), flo.to(DEV).requires_grad_(True)
    (ops.warp(xd, fd) * w.to(DEV)).sum().backward()
    close(xd.grad, xc.grad, atol=1e-5)
    close(fd.grad, fc.grad, atol=2e-4)


# ----------------------------------------------------------------------------- refine
def test_refine_golden(golden):
    g = golden("ops_refine")
    lt, lr, m, ce = (T(g[k]).to(DEV) for k in ("lt", "lr", "mask", "certs"))
    cfgs = {"full": (False, False, m, ce), "noM": (True, False, m, ce), "noP": (False, True, m, ce),
            "bare": (False, False, None, None)}
    for tag, (dM, dP, mask, certs) in cfgs.items():
        probs, label, maxp, trust = ops.refine_fused(lt, lr, mask, certs=certs, gamma=0.25, disable_M=dM, disable_P=dP)
        assert label.dtype == torch.int64 and torch.equal(label.cpu(), T(g[f"{tag}_label"])), tag
        close(probs, T(g[f"{tag}_probs"]), atol=3e-7)
        close(maxp, T(g[f"{tag}_maxprob"]), atol=3e-7)
    close(trust, T(g["trust"]), atol=0, rtol=1e-6)


@pytest.mark.parametrize("hw", [(64, 64), (512, 512), (37, 53)])
@pytest.mark.parametrize("use_logvar", [False, True])
def test_refine_bit_exact_vs_oracle(hw, use_logvar):
    """Pseudo-label map (int64) and every probability bit-identical to the oracle."""
    H, W = hw
    torch.manual_seed(H * 7 + W + int(use_logvar))
    lt, lr = torch.randn(2, 19, H, W) * 3, torch.randn(2, 19, H, W) * 3
    # force exact ties and static-class agreement on part of the image
    lt[:, :, :4] = lt[:, :1, :4]
    lr[0, :, 5:9] = lt[0, :, 5:9]
    mask = torch.rand(2, H, W) > 0.15
    certs = None if use_logvar else torch.rand(2, 1, H, W)
    logvar = torch.randn(2, 1, H, W) * 2 if use_logvar else None
    wp, wl, wmx, ws = oracle.refine(lt, lr, mask, certs=certs, logvar=logvar, gamma=0.25)
    gp, gl, gmx, gs = ops.refine_fused(lt.to(DEV), lr.to(DEV), mask.to(DEV), certs=None if certs is None else certs.to(DEV),
                                       logvar=None if logvar is None else logvar.to(DEV), gamma=0.25)
    assert torch.equal(gl.cpu(), wl)
    assert torch.equal(gs.cpu(), ws)
    assert torch.equal(gp.cpu(), wp)
    assert torch.equal(gmx.cpu(), wmx)


def test_refine_other_k_and_full_size_properties():
    torch.manual_seed(21)
    lt, lr = torch.randn(1, 7, 40, 40) * 2, torch.randn(1, 7, 40, 40) * 2
    wp, wl, _, _ = oracle.refine(lt, lr, None, static_classes=(0, 2, 5))
    gp, gl, _, _ = ops.refine_fused(lt.to(DEV), lr.to(DEV), None, static_classes=(0, 2, 5))
    assert torch.equal(gl.cpu(), wl) and torch.equal(gp.cpu(), wp)
    # BASELINE size (2 x 19 x 1024 x 1024): identical inputs => refined == softmax; mask off => eps = 0
    lt = torch.randn(2, 19, 1024, 1024, device=DEV) * 3
    sm = torch.softmax(lt, 1)
    p, lab, mx, s = ops.refine_fused(lt, lt.clone(), None, certs=torch.rand(2, 1, 1024, 1024, device=DEV))
    close(p, sm, atol=2e-6)
    assert (lab != sm.argmax(1)).float().mean().item() < 1e-5
    assert bool(((s > 0) & (s <= 1)).all())
    none_inside = torch.zeros(2, 1024, 1024, dtype=torch.bool, device=DEV)
    p2, _, _, _ = ops.refine_fused(lt, torch.randn_like(lt), none_inside)
    close(p2, sm, atol=2e-6)
    close(mx, p.max(1)[0], atol=0, rtol=0)


# ----------------------------------------------------------------------------- optimiser ops
def test_ema_and_adamw_vs_torch():
    torch.manual_seed(3)
    n = 100003
    live, ema = torch.randn(n), torch.randn(n)
    got = ops.ema_update_(ema.to(DEV), live.to(DEV), 0.999).cpu()
    assert torch.equal(got, ema * 0.999 + live * (1.0 - 0.999))
    # AdamW: 3 segments with their own lr / weight decay, 3 steps, vs torch.optim.AdamW on CPU
    ends, lrs, wds = [1000, 60000, n], [6e-4, 6e-5, 6e-4], [0.01, 0.01, 0.0]
    p0 = torch.randn(n)
    ref_params = [p0[a:b].clone().requires_grad_(True) for a, b in zip([0] + ends[:-1], ends)]
    opt = torch.optim.AdamW([{"params": [q], "lr": lr, "weight_decay": wd} for q, lr, wd in zip(ref_params, lrs, wds)],
                            betas=(0.9, 0.999), eps=1e-8)
    p, m, v = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        g = torch.randn(n)
        for q, a, b in zip(ref_params, [0] + ends[:-1], ends):
            q.grad = g[a:b].clone()
        opt.step()
        ops.adamw_step_(p, g.to(DEV), m, v, ends, lrs, wds, 0.9, 0.999, 1e-8, step)
    close(p, torch.cat([q.detach() for q in ref_params]), atol=1e-6, rtol=1e-5)
