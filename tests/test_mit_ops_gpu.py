"""GPU parity tests of the MiT / DAFormer operator kernels (depthwise conv, SR-attention, patch
embedding, LayerNorm) called through the C-ABI, against a plain PyTorch fp32 reference of the same
op (these are floating-point kernels; the tolerance is written in each test).

fp32 I/O: 1e-3 relative (north_star).  bf16 I/O: the result is compared with the fp32 reference
evaluated on the SAME bf16-rounded inputs; the tolerance is two bf16 ulps of the output scale
(2 * 2^-8), i.e. the rounding of the stored result, not of the arithmetic (which is fp32)."""
import pytest
import torch
import torch.nn.functional as F

from refign_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(got, want, rtol, atol, what=""):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    assert bool((err <= tol).all()), "%s: max abs err %.3e at |want| max %.3e" % (what, err.max().item(),
                                                                                  want.abs().max().item())


DW_SPECS = [  # B, H, W, C, dilation, gelu, bias
    (2, 16, 16, 64, 1, True, True),      # Mix-FFN shape family
    (1, 9, 13, 32, 1, True, True),       # ragged W (W % 4 != 0)
    (2, 24, 20, 128, 6, False, False),   # ASPP depthwise branches
    (1, 40, 40, 64, 12, False, False),
    (1, 40, 24, 16, 18, False, False),   # dilation close to the map size: most taps out of range
    (1, 5, 7, 8, 2, False, True),
    # shared-memory tile kernels (bf16, C % 64 == 0): ragged row / column tiles, both tile widths, several
    # channel blocks, with and without bias / GELU
    (1, 9, 13, 64, 1, True, True),
    (2, 20, 40, 128, 1, True, True),
    (1, 12, 96, 64, 1, True, True),
    (1, 16, 64, 192, 1, False, False),
    (2, 33, 70, 64, 1, False, True),
    (1, 50, 45, 64, 18, False, True),    # dilated lattice through the tile kernels (residue-class sub-images)
    (2, 37, 41, 128, 6, False, False),
]


@pytest.mark.parametrize("spec", DW_SPECS)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dwconv3x3_fwd_bwd(spec, dtype):
    B, H, W, C, dil, gelu, has_bias = spec
    torch.manual_seed(sum(spec[:5]))
    x = torch.randn(B, H, W, C, device=DEV).to(dtype)
    w = (torch.randn(C, 1, 3, 3, device=DEV) * 0.4).requires_grad_(True)
    b = (torch.randn(C, device=DEV) * 0.2).requires_grad_(True) if has_bias else None
    gy = torch.randn(B, H, W, C, device=DEV).to(dtype)
    xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    ref = F.conv2d(xr, w, b, padding=dil, dilation=dil, groups=C)
    if gelu:
        ref = F.gelu(ref)
    ref.backward(gy.float().permute(0, 3, 1, 2))
    gw_ref, gb_ref, gx_ref = w.grad.clone(), (b.grad.clone() if has_bias else None), xr.grad.permute(0, 2, 3, 1)
    w.grad = None
    if has_bias:
        b.grad = None
    xm = x.clone().requires_grad_(True)
    if gelu:
        out = ops.dwconv3x3_gelu(xm.view(B, H * W, C), H, W, w, b).view(B, H, W, C)
    else:
        out = ops.dwconv3x3_nhwc(xm.permute(0, 3, 1, 2), w, b, dil).permute(0, 2, 3, 1)
    assert out.dtype == dtype
    out.backward(gy)
    if dtype == torch.float32:
        rt, at = 1e-3, 1e-5
    else:
        rt, at = 2 * 2 ** -8, 2 * 2 ** -8
    _close(out, ref.permute(0, 2, 3, 1), rt, at, "forward")
    _close(xm.grad, gx_ref, rt, at * 4, "grad_input")
    # weight / bias gradients are fp32 sums over B*H*W terms; with bf16 I/O the GELU pre-pass stores a
    # bf16-rounded gradient, so the sums carry ~2^-9 relative noise per term
    scale = float(gw_ref.abs().max())
    _close(w.grad, gw_ref, rt, (1e-4 if dtype == torch.float32 else 2e-2) * max(scale, 1.0), "grad_weight")
    if has_bias:
        _close(b.grad, gb_ref, rt, (1e-4 if dtype == torch.float32 else 2e-2) * max(float(gb_ref.abs().max()), 1.0),
               "grad_bias")


def test_dwconv_rejects_bad_channels():
    x = torch.randn(1, 4, 4, 12, device=DEV)
    w = torch.randn(12, 1, 3, 3, device=DEV)
    with pytest.raises(RuntimeError):
        ops.dwconv3x3_nhwc(x.permute(0, 3, 1, 2), w, None, 1)


LN_SPECS = [  # B, N, C, with_branch, with_scale
    (2, 50, 64, False, False),
    (2, 33, 128, True, True),
    (3, 17, 320, True, False),
    (2, 40, 512, True, True),
    (2, 9, 32, True, True),       # generic (C % 64 != 0) path: mit_b0 widths
    (1, 7, 160, False, False),
    (2, 5, 256, True, False),
]


@pytest.mark.parametrize("spec", LN_SPECS)
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_add_layernorm_fwd_bwd(spec, mode):
    B, N, C, with_branch, with_scale = spec
    torch.manual_seed(B * 1000 + N * 10 + C)
    ln = torch.nn.LayerNorm(C, eps=1e-6).to(DEV)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5)
        ln.bias.uniform_(-0.5, 0.5)
    bdt = torch.bfloat16 if mode == "bf16" else torch.float32
    ydt = bdt
    x = (torch.randn(B, N, C, device=DEV) * 2 + 0.3).requires_grad_(True)
    br = torch.randn(B, N, C, device=DEV).to(bdt).requires_grad_(True) if with_branch else None
    sc = (torch.rand(B, device=DEV) > 0.3).float() / 0.7 if with_scale else None
    gy = torch.randn(B, N, C, device=DEV).to(ydt)
    gxn = torch.randn(B, N, C, device=DEV)
    # reference in fp32 on the same (rounded) inputs
    xr = x.detach().clone().requires_grad_(True)
    brr = br.detach().float().clone().requires_grad_(True) if with_branch else None
    xn_r = xr if not with_branch else xr + (brr if sc is None else brr * sc.view(-1, 1, 1))
    y_r = F.layer_norm(xn_r, (C,), ln.weight, ln.bias, ln.eps)
    loss_r = (y_r * gy.float()).sum() + ((xn_r * gxn).sum() if with_branch else 0)
    loss_r.backward()
    ref_g = (ln.weight.grad.clone(), ln.bias.grad.clone())
    ln.weight.grad = ln.bias.grad = None
    if with_branch:
        xn, y = ops.add_layer_norm(x, br, sc, ln, out_dtype=ydt)
        ((y * gy).sum() + (xn * gxn).sum()).backward()
        _close(xn, xn_r, 1e-6, 1e-6, "xn")
    else:
        y = ops.layer_norm(x, ln, out_dtype=ydt)
        (y * gy).sum().backward()
    assert y.dtype == ydt
    rt, at = (1e-3, 1e-4) if mode == "fp32" else (2 * 2 ** -8, 2 * 2 ** -8)
    _close(y, y_r, rt, at, "y")
    _close(x.grad, xr.grad, 1e-3, 1e-3 if mode == "fp32" else 2e-2, "dx")
    if with_branch:
        _close(br.grad, brr.grad, rt, 1e-3 if mode == "fp32" else 3e-2, "dbranch")
    _close(ln.weight.grad, ref_g[0], 1e-3, 1e-3 * max(1.0, float(ref_g[0].abs().max())), "dgamma")
    _close(ln.bias.grad, ref_g[1], 1e-3, 1e-3 * max(1.0, float(ref_g[1].abs().max())), "dbeta")


AT_SPECS = [  # B, heads, N, M
    (1, 1, 128, 128),       # one tile, one chunk
    (2, 2, 256, 256),       # 512x512 input shapes: M = 256
    (1, 5, 200, 256),       # ragged query tile
    (2, 1, 384, 1024),      # 1024x1024 input shapes: M = 1024 (8 chunks)
    (1, 8, 1024, 1024),     # stage-4 shape at 1024x1024
    (1, 2, 130, 72),        # ragged key chunk (masking) + ragged query tile
    (1, 1, 64, 320),        # N < tile, M not a multiple of the chunk
]


def _attention_ref(q, kv, heads, scale):
    B, N, C = q.shape
    M = kv.shape[1]
    d = C // heads
    q4 = q.float().view(B, N, heads, d).transpose(1, 2)
    k4 = kv[..., :C].float().reshape(B, M, heads, d).transpose(1, 2)
    v4 = kv[..., C:].float().reshape(B, M, heads, d).transpose(1, 2)
    s = (q4 @ k4.transpose(-2, -1)) * scale
    p = torch.softmax(s, dim=-1)
    return (p @ v4).transpose(1, 2).reshape(B, N, C), torch.logsumexp(s, dim=-1)


@pytest.mark.parametrize("spec", AT_SPECS)
def test_sr_attention_fwd(spec):
    """bf16 operands, fp32 accumulation; P is rounded to bf16 before the PV product (as every
    flash-attention kernel does) and the output is stored in bf16 -> tolerance 2^-7 relative to the
    output scale; the fp32 log-sum-exp is held to 1e-3."""
    B, heads, N, M = spec
    torch.manual_seed(B + heads * 10 + N + M)
    C = heads * 64
    q = (torch.randn(B, N, C, device=DEV) * 1.5).bfloat16()
    kv = (torch.randn(B, M, 2 * C, device=DEV) * 1.5).bfloat16()
    scale = 0.125
    out, lse = ops.sr_attention_fwd(q, kv, heads, scale, want_lse=True)
    ref, lse_ref = _attention_ref(q, kv, heads, scale)
    assert out.dtype == torch.bfloat16 and out.shape == q.shape
    _close(lse, lse_ref, 1e-3, 1e-3, "lse")
    _close(out, ref, 2 ** -7, 2 ** -7 * float(ref.abs().max()), "out")


@pytest.mark.parametrize("spec", AT_SPECS)
def test_sr_attention_bwd(spec):
    """Fused backward (P recomputed from q, k and the fp32 LSE) against autograd of the fp32 reference on
    the same bf16 inputs.  dS is rounded to bf16 before its two products and the gradients are stored in
    bf16: tolerance 2^-6 relative to the gradient scale."""
    B, heads, N, M = spec
    torch.manual_seed(B + heads * 10 + N + M + 1)
    C = heads * 64
    q = torch.randn(B, N, C, device=DEV).bfloat16().requires_grad_(True)
    kv = torch.randn(B, M, 2 * C, device=DEV).bfloat16().requires_grad_(True)
    go = torch.randn(B, N, C, device=DEV).bfloat16()
    out = ops.sr_attention(q, kv, heads, 0.125)
    out.backward(go)
    gq, gkv = q.grad.clone(), kv.grad.clone()
    q.grad = kv.grad = None
    qr = q.detach().float().requires_grad_(True)
    kvr = kv.detach().float().requires_grad_(True)
    ref, _ = _attention_ref(qr, kvr, heads, 0.125)
    ref.backward(go.float())
    _close(gq, qr.grad, 2 ** -6, 2 ** -6 * float(qr.grad.abs().max()), "dq")
    _close(gkv[..., :C], kvr.grad[..., :C], 2 ** -6, 2 ** -6 * float(kvr.grad[..., :C].abs().max()), "dk")
    _close(gkv[..., C:], kvr.grad[..., C:], 2 ** -6, 2 ** -6 * float(kvr.grad[..., C:].abs().max()), "dv")


@pytest.mark.parametrize("hd", [64, 32])
@pytest.mark.parametrize("spec", AT_SPECS + [(2, 5, 1024, 256), (1, 2, 4096, 1024)])
def test_sr_attention_fp32_mode(spec, hd):
    """fp32 parity mode (precision='fp32'): exact FFMA kernels through the C-ABI, forward + backward against a
    float64 evaluation of the reference formulation (mix_transformer.py:150-160).  Tolerance 1e-5 relative to each
    tensor's scale (summation-order noise of fp32) -- two orders inside north_star's 1e-3."""
    B, heads, N, M = spec
    torch.manual_seed(B + heads * 10 + N + M + 7)
    C = heads * hd
    q = (torch.randn(B, N, C, device=DEV) * 1.5).requires_grad_(True)
    kv = (torch.randn(B, M, 2 * C, device=DEV) * 1.5).requires_grad_(True)
    go = torch.randn(B, N, C, device=DEV)
    out, lse = ops.sr_attention_fwd(q.detach(), kv.detach(), heads, 0.125, want_lse=True)
    o2 = ops.sr_attention(q, kv, heads, 0.125)
    assert o2.dtype == torch.float32 and torch.equal(o2.detach(), out)
    o2.backward(go)
    qr = q.detach().double().requires_grad_(True)
    kvr = kv.detach().double().requires_grad_(True)
    d = hd
    q4 = qr.view(B, N, heads, d).transpose(1, 2)
    k4 = kvr[..., :C].reshape(B, M, heads, d).transpose(1, 2)
    v4 = kvr[..., C:].reshape(B, M, heads, d).transpose(1, 2)
    sc = (q4 @ k4.transpose(-2, -1)) * 0.125
    ref = (torch.softmax(sc, -1) @ v4).transpose(1, 2).reshape(B, N, C)
    ref.backward(go.double())
    _close(lse, torch.logsumexp(sc, -1).float(), 1e-5, 1e-5, "lse")
    _close(out, ref.float(), 1e-5, 1e-5 * float(ref.abs().max()), "out")
    _close(q.grad, qr.grad.float(), 1e-5, 2e-5 * float(qr.grad.abs().max()), "dq")
    _close(kv.grad[..., :C], kvr.grad[..., :C].float(), 1e-5, 2e-5 * float(kvr.grad[..., :C].abs().max()), "dk")
    _close(kv.grad[..., C:], kvr.grad[..., C:].float(), 1e-5, 2e-5 * float(kvr.grad[..., C:].abs().max()), "dv")


def test_sr_attention_has_no_library_fallback():
    """Unsupported operands raise (north_star: no multi-backend dispatch, no fallback)."""
    q = torch.randn(1, 128, 64, device=DEV).half()
    kv = torch.randn(1, 128, 128, device=DEV).half()
    with pytest.raises(RuntimeError):
        ops.sr_attention(q, kv, 1, 0.125)
    with pytest.raises(RuntimeError):
        ops.sr_attention(q.float().cpu(), kv.float().cpu(), 1, 0.125)


def test_graphed_train_step_matches_eager():
    """The CUDA-graph replay of the train step (enable_cuda_graphs) must follow the same trajectory as
    the eager step: 6 steps, mit_b0, 128x128, randomness off (drop-path / dropout 0, no jitter / blur).
    Tolerance 2e-3 on the parameters: the fp32 atomics of the weight-gradient kernels make both runs
    order-dependent at the 1e-6 level and Adam amplifies that during its first steps."""
    import bench
    import refign_b200 as P

    def build():
        torch.manual_seed(0)
        dims = P.MixVisionTransformer.arch_settings['mit_b0']['embed_dims']
        m = P.DomainAdaptationSegmentationModel(
            optimizer_init={'class_path': 'torch.optim.AdamW', 'init_args': {'lr': 6e-4, 'weight_decay': 0.01, 'eps': 1e-2}},
            lr_scheduler_init=bench.SCH, backbone=P.MixVisionTransformer('mit_b0', drop_path_rate=0.0),
            head=P.DAFormerHead(dims, [0, 1, 2, 3], 19, 'multiple_select', dropout_ratio=0.0),
            loss=P.PixelWeightedCrossEntropyLoss(), alignment_backbone=P.VGG('vgg16', out_indices=[2, 3, 4]),
            alignment_head=P.UAWarpCHead(in_index=[0, 1], input_transform='multiple_select', estimate_uncertainty=True),
            backbone_lr_factor=0.1, enable_fdist=True, use_refign=True, adapt_to_ref=False, color_jitter_p=1.1,
            blur=False, precision='bf16').to(DEV).train()
        with torch.no_grad():   # a distinct ImageNet copy so that the feature distance has a gradient
            g = torch.Generator(device=DEV).manual_seed(1)
            for p_ in m.imnet_backbone.parameters():
                p_.add_(0.02 * torch.randn(p_.shape, device=DEV, generator=g))
        m.setup_runtime()
        return m

    batch = bench.synth_batch(128, 2, 5, torch.device(DEV))
    runs = []
    # reference trajectory: eager, one stream, the reference's two source backward passes; then the graphed
    # step with the concurrent branches and the single source backward (the defaults)
    for graphed in (False, True):
        m = build()
        if graphed:
            m.enable_cuda_graphs(warmup=2)
        else:
            m.concurrent_branches = False
            m.fuse_source_backward = False
            m.fused_loss = False     # library F.interpolate + cross-entropy, as the reference runs it
        torch.manual_seed(123)
        for i in range(6):
            m.training_step(batch, i)
        torch.cuda.synchronize()
        runs.append((m._rt['live'].data.clone(), m._rt['ema'].data.clone(),
                     {k: float(v) for k, v in m._logged.items()}))
        del m
    (p0, e0, l0), (p1, e1, l1) = runs
    for k in l0:
        assert abs(l0[k] - l1[k]) <= 2e-2 * max(1.0, abs(l0[k])), (k, l0[k], l1[k])
    _close(p1, p0, 2e-3, 2e-3, "parameters")
    _close(e1, e0, 2e-3, 2e-3, "ema parameters")


@pytest.mark.parametrize("shape", [(2, 19, 64, 64, 256, 256), (1, 19, 24, 20, 100, 84), (2, 7, 16, 16, 16, 16),
                                   (1, 19, 8, 8, 31, 33), (1, 32, 5, 3, 40, 12)])
@pytest.mark.parametrize("weighted", [False, True])
def test_upsample_cross_entropy_fwd_bwd(shape, weighted):
    """Fused bilinear up-sampling + pixel-weighted cross-entropy (ignore_index 255, mean over ALL pixels) vs
    the reference formulation F.interpolate(align_corners=False) + PixelWeightedCrossEntropyLoss
    (models/losses.py:10-22) evaluated by torch on the CPU in fp32.  Loss 1e-5 relative; gradient 1e-4 relative
    + 1e-5 of its largest entry (fp32 on both sides, different summation order)."""
    import refign_b200 as P
    B, K, h, w, H, W = shape
    torch.manual_seed(sum(shape) + int(weighted))
    low = (3.0 * torch.randn(B, K, h, w)).requires_grad_(True)
    tgt = torch.randint(0, K, (B, H, W))
    tgt[torch.rand(B, H, W) < 0.1] = 255
    pw = torch.rand(B, H, W) if weighted else None
    up = F.interpolate(low.float(), (H, W), mode='bilinear', align_corners=False)
    want = P.PixelWeightedCrossEntropyLoss()(up, tgt, pixel_weight=pw)
    (want * 1.7).backward()
    g_want = low.grad.clone()
    lg = low.detach().to(DEV).requires_grad_(True)
    got = ops.upsample_cross_entropy(lg, tgt.to(DEV), None if pw is None else pw.to(DEV), 255)
    assert got.shape == () and abs(float(got) - float(want)) <= 1e-5 * max(1.0, abs(float(want))), (float(got), float(want))
    (got * 1.7).backward()
    _close(lg.grad.cpu(), g_want, 1e-4, 1e-5 * float(g_want.abs().max()), "grad_logits")
    # all pixels ignored: loss 0 and a zero gradient
    lg2 = low.detach().to(DEV).requires_grad_(True)
    z = ops.upsample_cross_entropy(lg2, torch.full((B, H, W), 255, device=DEV), None, 255)
    z.backward()
    assert float(z) == 0.0 and float(lg2.grad.abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        ops.upsample_cross_entropy(lg, tgt.to(DEV)[0], None, 255)   # 2-D target


def test_upsample_cross_entropy_golden():
    """The fused loss tail against the values and gradients the reference's own F.interpolate +
    models.losses.PixelWeightedCrossEntropyLoss produced (tests/golden/ops_upsample_ce.npz, generated by
    tests/golden/make_golden_loss.py in the build container).  Same tolerances as the torch comparison above."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ops_upsample_ce.npz"))
    for i in range(int(g["ncases"])):
        lg = torch.from_numpy(g[f"c{i}_logits"]).to(DEV).requires_grad_(True)
        tg = torch.from_numpy(g[f"c{i}_target"]).to(DEV)
        w = torch.from_numpy(g[f"c{i}_weight"])
        got = ops.upsample_cross_entropy(lg, tg, w.to(DEV) if w.numel() else None, 255)
        want = float(g[f"c{i}_loss"])
        assert abs(float(got.detach()) - want) <= 1e-5 * max(1.0, abs(want)), (i, float(got.detach()), want)
        got.backward()
        g_want = torch.from_numpy(g[f"c{i}_grad"])
        _close(lg.grad.cpu(), g_want, 1e-4, 1e-5 * float(g_want.abs().max()), "grad_logits case %d" % i)


@pytest.mark.parametrize("shape", [(4, 19, 64, 64, 256, 256), (1, 3, 24, 20, 100, 84), (2, 1, 16, 16, 16, 16)])
def test_upsample_bilinear_f32(shape):
    """Teacher-logit up-sampling vs F.interpolate(mode='bilinear', align_corners=False), fp32: 1e-5."""
    B, C, h, w, H, W = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(B, C, h, w, device=DEV)
    _close(ops.upsample_bilinear(x, (H, W)), F.interpolate(x, (H, W), mode='bilinear', align_corners=False), 1e-5, 1e-5,
           "upsampled")


@pytest.mark.parametrize("spec", [(2, 64, 64, 64), (1, 64, 100, 76), (2, 32, 40, 56), (1, 64, 7, 9)])
def test_patch_embed_ln_fwd_bwd(spec):
    """Fused 7x7/s4 conv (3 -> C) + LayerNorm(eps 1e-5) vs torch conv2d + layer_norm in fp32 (TF32 off).
    fp32 arithmetic on both sides: 1e-3 relative (north_star), observed ~1e-5."""
    B, C, H, W = spec
    torch.manual_seed(sum(spec))
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        x = torch.randn(B, 3, H, W, device=DEV)
        conv = torch.nn.Conv2d(3, C, 7, 4, 3).to(DEV)
        ln = torch.nn.LayerNorm(C).to(DEV)
        with torch.no_grad():
            ln.weight.uniform_(0.5, 1.5)
            ln.bias.uniform_(-0.5, 0.5)
        ref = F.layer_norm(conv(x).flatten(2).transpose(1, 2), (C,), ln.weight, ln.bias, ln.eps)
        gy = torch.randn_like(ref)
        ref.backward(gy)
        want = [p.grad.clone() for p in (conv.weight, conv.bias, ln.weight, ln.bias)]
        for p in (conv.weight, conv.bias, ln.weight, ln.bias):
            p.grad = None
        y, Ho, Wo = ops.patch_embed_ln(x, conv.weight, conv.bias, ln.weight, ln.bias, ln.eps)
        assert (Ho, Wo) == ((H - 1) // 4 + 1, (W - 1) // 4 + 1) and y.shape == ref.shape
        _close(y, ref, 1e-3, 1e-4, "tokens")
        y.backward(gy)
        for name, p, w_ in zip(("dW", "db", "dgamma", "dbeta"), (conv.weight, conv.bias, ln.weight, ln.bias), want):
            _close(p.grad, w_, 1e-3, 1e-3 * max(1.0, float(w_.abs().max())), name)
        with torch.no_grad():
            y2, _, _ = ops.patch_embed_ln(x, conv.weight, conv.bias, ln.weight, ln.bias, ln.eps)
        assert torch.equal(y2, y.detach())
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("spec", [(2, 16, 12, 64, True), (3, 9, 7, 256, False), (2, 32, 32, 1024, True)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_batch_norm_relu_fwd_bwd(spec, dtype):
    """Training-mode BatchNorm (+ ReLU) on channels-last activations vs nn.BatchNorm2d + ReLU in fp32 on the
    same (rounded) inputs, including the running-statistics update.  fp32: 1e-3 relative; bf16 I/O: two
    bf16 ulps of the output scale."""
    B, H, W, C, relu = spec
    torch.manual_seed(sum(spec[:4]))
    x = (torch.randn(B, C, H, W, device=DEV) * 1.7 + 0.4).to(dtype).contiguous(memory_format=torch.channels_last)
    gy = torch.randn(B, C, H, W, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last)
    ref_bn = torch.nn.BatchNorm2d(C).to(DEV).train()
    with torch.no_grad():
        ref_bn.weight.uniform_(0.5, 1.5)
        ref_bn.bias.uniform_(-0.5, 0.5)
    import copy
    my_bn = copy.deepcopy(ref_bn)
    xr = x.detach().clone().float().requires_grad_(True)
    yr = ref_bn(xr)
    if relu:
        yr = torch.relu(yr)
    yr.backward(gy.float())
    xm = x.detach().clone().requires_grad_(True)
    ym = ops.batch_norm_act(xm, my_bn, relu)
    assert ym.dtype == dtype and ym.shape == x.shape
    ym.backward(gy)
    rt, at = (1e-3, 1e-4) if dtype == torch.float32 else (2 * 2 ** -8, 2 * 2 ** -8)
    _close(ym, yr, rt, at * max(1.0, float(yr.abs().max())), "y")
    _close(xm.grad, xr.grad, rt, (1e-4 if dtype == torch.float32 else 3e-2) * max(1.0, float(xr.grad.abs().max())), "dx")
    _close(my_bn.weight.grad, ref_bn.weight.grad, 1e-3, 2e-3 * max(1.0, float(ref_bn.weight.grad.abs().max())), "dgamma")
    _close(my_bn.bias.grad, ref_bn.bias.grad, 1e-3, 2e-3 * max(1.0, float(ref_bn.bias.grad.abs().max())), "dbeta")
    _close(my_bn.running_mean, ref_bn.running_mean, 1e-4, 1e-5, "running_mean")
    _close(my_bn.running_var, ref_bn.running_var, 1e-4, 1e-5, "running_var")
    assert int(my_bn.num_batches_tracked) == 1


def test_direct_param_grad_accumulation_matches_autograd():
    """Backward kernels that accumulate straight into a bound ``.grad`` (runtime.FlatParams marks the
    parameters ``_rf_direct_grad``) must give what autograd's AccumulateGrad gives over two backward
    passes, for the Linear (shadow bf16 GEMM + colsum), LayerNorm / add+LayerNorm and depthwise-conv paths."""
    import torch.nn as nn
    torch.manual_seed(3)
    B, H, W, C = 2, 16, 16, 64
    lin, ln, ln2 = nn.Linear(C, 2 * C).to(DEV), nn.LayerNorm(C).to(DEV), nn.LayerNorm(C).to(DEV)
    dw_w = nn.Parameter(torch.randn(2 * C, 1, 3, 3, device=DEV) * 0.3)
    dw_b = nn.Parameter(torch.randn(2 * C, device=DEV) * 0.1)
    params = [lin.weight, lin.bias, ln.weight, ln.bias, ln2.weight, ln2.bias, dw_w, dw_b]
    for p in (lin.weight, lin.bias):
        p._rf_bf16 = p.detach().to(torch.bfloat16)
    xs = [torch.randn(B, H * W, C, device=DEV) for _ in range(2)]

    def run(direct):
        for p in params:
            p.grad = torch.full_like(p, 0.25)      # pre-existing content must be added to, not overwritten
            p._rf_direct_grad = direct
        for x in xs:
            with torch.autocast('cuda', dtype=torch.bfloat16):
                xn, y = ops.add_layer_norm(x, x * 0.5, None, ln)
                h = ops.linear(y, lin.weight, lin.bias)
                h = ops.dwconv3x3_gelu(h, H, W, dw_w, dw_b)
                z = ops.layer_norm(xn, ln2)
            (h.float().square().mean() + z.float().sum() * 1e-3).backward()
        return [p.grad.clone() for p in params]

    want = run(False)
    got = run(True)
    for p in params:
        p._rf_direct_grad = False
    for g, w_, name in zip(got, want, ["lin.w", "lin.b", "ln.w", "ln.b", "ln2.w", "ln2.b", "dw.w", "dw.b"]):
        _close(g, w_, 2e-3, 2e-3 * max(1.0, float(w_.abs().max())), name)


@pytest.mark.parametrize("spec", [(2, 32, 32, 64, 8), (1, 16, 24, 128, 4), (2, 8, 8, 320, 2)])
def test_sr_conv_patch_gemm_matches_conv(spec):
    """Spatial-reduction conv (kernel == stride) as space-to-depth + GEMM vs nn.Conv2d in fp32 on the same
    bf16-rounded inputs / weights: forward, input / weight / bias gradients (direct accumulation too)."""
    import torch.nn as nn
    B, H, W, C, s = spec
    torch.manual_seed(sum(spec))
    conv = nn.Conv2d(C, C, kernel_size=s, stride=s).to(DEV)
    with torch.no_grad():
        conv.weight.mul_(3.0)
    conv.weight._rf_bf16 = conv.weight.detach().to(torch.bfloat16)
    conv.bias._rf_bf16 = conv.bias.detach().to(torch.bfloat16)
    x = torch.randn(B, H * W, C, device=DEV).to(torch.bfloat16)
    gy = torch.randn(B, (H // s) * (W // s), C, device=DEV).to(torch.bfloat16)
    xr = x.float().view(B, H, W, C).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wr = conv.weight._rf_bf16.float().requires_grad_(True)
    br = conv.bias._rf_bf16.float().requires_grad_(True)
    ref = F.conv2d(xr, wr, br, stride=s)
    ref.backward(gy.float().view(B, H // s, W // s, C).permute(0, 3, 1, 2))
    for direct in (False, True):
        conv.weight.grad = torch.zeros_like(conv.weight)
        conv.bias.grad = torch.zeros_like(conv.bias)
        conv.weight._rf_direct_grad = conv.bias._rf_direct_grad = direct
        xm = x.clone().requires_grad_(True)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            out = ops.sr_conv(xm, H, W, conv)
        assert out.dtype == torch.bfloat16 and out.shape == gy.shape
        out.backward(gy)
        tol = 2 * 2 ** -8
        _close(out, ref.permute(0, 2, 3, 1).reshape(gy.shape), tol, tol * float(ref.abs().max()), "forward")
        _close(xm.grad, xr.grad.permute(0, 2, 3, 1).reshape(x.shape), tol, tol * float(xr.grad.abs().max()), "grad_input")
        _close(conv.weight.grad, wr.grad, tol, tol * float(wr.grad.abs().max()), "grad_weight")
        _close(conv.bias.grad, br.grad, tol, tol * float(br.grad.abs().max()), "grad_bias")
    # a refresh of the shadow must be followed by the derived (permuted) weight
    with torch.no_grad():
        conv.weight._rf_bf16.mul_(0.5)
    ops.refresh_derived(None)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        out2 = ops.sr_conv(x, H, W, conv)
    ref2 = F.conv2d(xr.detach(), conv.weight._rf_bf16.float(), br.detach(), stride=s)
    _close(out2, ref2.permute(0, 2, 3, 1).reshape(gy.shape), 2 * 2 ** -8, 2 * 2 ** -8 * float(ref2.abs().max()), "refreshed")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("spec", [(2, 64, 24, 40, True, "relu"), (3, 32, 7, 7, False, "leaky"), (1, 19, 16, 16, True, None),
                                  (2, 128, 8, 8, False, "relu")])
def test_conv_bias_act_matches_library(spec, dtype):
    """Library conv + fused in-place bias/activation sweep (vector and scalar paths, both memory formats)
    vs conv2d + bias + activation."""
    import torch.nn as nn
    B, C, H, W, channels_last, actname = spec
    torch.manual_seed(sum(spec[:4]))
    act = {"relu": nn.ReLU(), "leaky": nn.LeakyReLU(0.1), None: None}[actname]
    x = torch.randn(B, 16, H, W, device=DEV).to(dtype)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    w = (torch.randn(C, 16, 3, 3, device=DEV) * 0.2).to(dtype)
    b = torch.randn(C, device=DEV)
    ref = F.conv2d(x.float(), w.float(), b, padding=1)
    if act is not None:
        ref = act(ref)
    with torch.no_grad():
        got = ops.conv_bias_act(x, w, b, 1, 1, 1, 1, act)
    assert got.dtype == dtype
    tol = 1e-3 if dtype == torch.float32 else 2 * 2 ** -8
    _close(got, ref, tol, tol * float(ref.abs().max()), "conv_bias_act")


@pytest.mark.parametrize("spec", [(2, 32, 32, [(32, 32, 64), (16, 16, 64), (8, 8, 64), (4, 4, 64)]),
                                  (1, 24, 40, [(24, 40, 32), (12, 20, 16), (5, 7, 8)]),
                                  (2, 16, 16, [(8, 8, 256), (16, 16, 8)])])
def test_upsample_concat_fwd_bwd(spec):
    """Fused bilinear resize + concat vs F.interpolate(align_corners=False) + cat in fp32 on the same bf16
    inputs; gradients of every source vs autograd."""
    B, H, W, srcs = spec
    torch.manual_seed(H + W)
    feats = [torch.randn(B, h * w, E, device=DEV).to(torch.bfloat16) for h, w, E in srcs]
    sizes = [(h, w) for h, w, _ in srcs]
    refs = [f.float().clone().requires_grad_(True) for f in feats]
    maps = []
    for r, (h, w, E) in zip(refs, srcs):
        m = r.view(B, h, w, E).permute(0, 3, 1, 2)
        if (h, w) != (H, W):
            m = F.interpolate(m, size=(H, W), mode="bilinear", align_corners=False)
        maps.append(m)
    ref = torch.cat(maps, 1)
    gy = torch.randn_like(ref).to(torch.bfloat16)
    ref.backward(gy.float())
    ins = [f.clone().requires_grad_(True) for f in feats]
    out = ops.upsample_concat(ins, sizes, (H, W))
    assert out.shape == ref.shape and out.dtype == torch.bfloat16
    assert out.is_contiguous(memory_format=torch.channels_last)
    out.backward(gy)
    tol = 2 * 2 ** -8
    _close(out, ref, tol, tol * float(ref.abs().max()), "forward")
    for i, (a, r) in enumerate(zip(ins, refs)):
        _close(a.grad, r.grad, tol, tol * float(r.grad.abs().max()), "grad_src%d" % i)
