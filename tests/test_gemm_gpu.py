"""tcgen05 bf16 GEMM (csrc/gemm_bf16.cu through ops.gemm_bf16) against a plain PyTorch fp32 reference of the same
product on the same bf16 operands: every operand-layout variant the Linear layers use (forward: K-major x K-major;
input gradient: K-major x MN-major; weight gradient: MN-major x MN-major, fp32 accumulate with split contraction),
ragged M / N / K, the MiT-B5 layer shapes, and ops.linear's autograd (forward, dx, dW, db) against F.linear.
Tolerances: fp32 accumulation of bf16 products -> 2e-3 of the output scale for bf16 outputs (one bf16 rounding),
1e-4 for fp32 outputs."""
import pytest
import torch
import torch.nn.functional as F

from refign_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [  # M, N, K
    (128, 128, 64), (256, 64, 64), (384, 320, 128), (1000, 136, 72), (8192, 1280, 320), (8192, 320, 1280),
    (2048, 512, 2048), (4096, 64, 64), (520, 2048, 512), (131072, 64, 64),
]


def _close(got, want, tol, what):
    scale = float(want.abs().max()) + 1e-6
    err = float((got.float() - want).abs().max())
    assert err <= tol * scale, (what, err, scale)


@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_forward_layout(shape):
    M, N, K = shape
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV).bfloat16()
    b = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV)
    want = a.float() @ b.float().t() + bias
    _close(ops.gemm_bf16(a, b, bias), want, 4e-3, "bf16 out + bias")
    _close(ops.gemm_bf16(a, b, None, out_dtype=torch.float32), want - bias, 1e-4, "f32 out")


@pytest.mark.parametrize("shape", SHAPES[:9])
def test_gemm_dgrad_and_wgrad_layouts(shape):
    T, N, K = shape            # y [T,N] = x [T,K] W[N,K]^T
    torch.manual_seed(T + N + K + 1)
    x = torch.randn(T, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    dy = torch.randn(T, N, device=DEV).bfloat16()
    # dx = dy W  (B = W read MN-major)
    _close(ops.gemm_bf16(dy, w, b_mn_major=True), dy.float() @ w.float(), 4e-3, "dx")
    # dW += dy^T x  (both MN-major, fp32 accumulate onto an existing gradient, split contraction)
    g0 = torch.randn(N, K, device=DEV)
    g = g0.clone()
    ops.gemm_bf16(dy, x, out=g, a_mn_major=True, b_mn_major=True, accumulate=True)
    _close(g, g0 + dy.float().t() @ x.float(), 2e-4, "dW accumulate")


def test_linear_autograd_matches_library():
    """ops.linear with bf16 shadow weights under autocast: forward, dx, dW (accumulated into the bound fp32 gradient), db."""
    torch.manual_seed(3)
    lin = torch.nn.Linear(320, 1280).to(DEV)
    lin.weight._rf_bf16 = lin.weight.detach().bfloat16()
    lin.bias._rf_bf16 = lin.bias.detach().bfloat16()
    x = torch.randn(2, 1024, 320, device=DEV).bfloat16().requires_grad_(True)
    gy = torch.randn(2, 1024, 1280, device=DEV).bfloat16()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        y = ops.linear(x, lin.weight, lin.bias)
    assert y.dtype == torch.bfloat16
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    wr = lin.weight.detach().bfloat16().float().requires_grad_(True)
    br = lin.bias.detach().clone().requires_grad_(True)
    yr = F.linear(xr, wr, br)
    yr.backward(gy.float())
    _close(y, yr, 4e-3, "y")
    _close(x.grad, xr.grad, 4e-3, "dx")
    _close(lin.weight.grad, wr.grad, 2e-4, "dW")
    _close(lin.bias.grad, br.grad, 1e-3, "db")


CONV_SHAPES = [  # B, H, W, Cin, Cout
    (1, 8, 16, 64, 64), (2, 24, 40, 64, 128), (1, 19, 21, 128, 72), (2, 32, 32, 1024, 256), (1, 64, 64, 256, 512),
    (1, 16, 48, 72, 64),
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv3x3_forward_bias_act(shape):
    """Implicit-GEMM 3x3 convolution (+ bias + ReLU / LeakyReLU) vs F.conv2d in fp32 on the same bf16 operands."""
    B, H, W, Ci, Co = shape
    torch.manual_seed(sum(shape))
    x = torch.randn(B, H, W, Ci, device=DEV).bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, device=DEV) / (3 * Ci ** 0.5)).bfloat16()
    b = torch.randn(Co, device=DEV)
    w_cl = w.permute(0, 2, 3, 1).contiguous()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1)
    for act, slope, fn in ((0, 0.0, lambda t: t), (1, 0.0, torch.relu), (2, 0.1, lambda t: F.leaky_relu(t, 0.1))):
        got = ops.conv3x3_nhwc_raw(x, w_cl, b, act=act, slope=slope)
        _close(got.permute(0, 3, 1, 2), fn(ref), 4e-3, "conv act %d" % act)
    got32 = ops.conv3x3_nhwc_raw(x, w_cl, None, out_dtype=torch.float32)
    _close(got32.permute(0, 3, 1, 2), ref - b.view(1, -1, 1, 1), 1e-4, "conv f32")


@pytest.mark.parametrize("shape", CONV_SHAPES[:5])
def test_conv3x3_autograd(shape):
    """_Conv3x3: forward, input gradient (flipped / transposed filter through the same kernel) and weight gradient
    (accumulated into the bound fp32 gradient) vs autograd of F.conv2d."""
    B, H, W, Ci, Co = shape
    torch.manual_seed(sum(shape) + 1)
    conv = torch.nn.Conv2d(Ci, Co, 3, padding=1, bias=False).to(DEV)
    conv.weight._rf_bf16 = conv.weight.detach().bfloat16()
    x = torch.randn(B, Ci, H, W, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gy = torch.randn(B, Co, H, W, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    assert ops.conv3x3_supported(x, conv)
    y = ops.conv3x3_train(x, conv)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    wr = conv.weight.detach().bfloat16().float().requires_grad_(True)
    yr = F.conv2d(xr, wr, padding=1)
    yr.backward(gy.float())
    _close(y, yr, 4e-3, "y")
    _close(x.grad, xr.grad, 4e-3, "dx")
    _close(conv.weight.grad, wr.grad, 3e-4, "dW")


@pytest.mark.parametrize("spec", [  # B, Cin, Cout, H, W, kernel, dilation, act
    (2, 84, 128, 32, 32, 3, 1, None), (2, 86, 128, 24, 40, 3, 1, "leaky"), (1, 32, 2, 64, 64, 3, 1, None),
    (2, 128, 128, 32, 32, 3, 2, "leaky"), (1, 128, 96, 40, 24, 3, 16, "leaky"), (2, 16, 1, 16, 16, 3, 1, None),
    (2, 128, 96, 32, 32, 1, 1, None), (1, 96, 32, 24, 40, 1, 1, None), (1, 41, 32, 16, 16, 3, 4, "relu"),
])
def test_frozen_decoder_convs_on_the_implicit_gemm(spec):
    """conv_bias_act on the shapes of the alignment decoders (reference models/modules.py:395-477): channel counts that
    are not multiples of 8 (zero-padded), dilated 3x3 filters of the RefinementModule (padding = dilation through TMA's
    out-of-bounds zero fill), 1x1 skip convolutions on the GEMM -- against conv2d in fp32 on the same bf16 operands."""
    import torch.nn as nn
    B, Ci, Co, H, W, k, dil, actname = spec
    torch.manual_seed(Ci * Co + H + dil)
    act = {"relu": nn.ReLU(), "leaky": nn.LeakyReLU(0.1), None: None}[actname]
    x = torch.randn(B, Ci, H, W, device=DEV).bfloat16()
    w = (torch.randn(Co, Ci, k, k, device=DEV) / (k * Ci ** 0.5))
    b = torch.randn(Co, device=DEV)
    pad = dil if k == 3 else 0
    want = F.conv2d(x.float(), w.bfloat16().float(), b, padding=pad, dilation=dil)
    if act is not None:
        want = act(want)
    timer = ops.KernelTimer()
    ops.set_timer(timer)
    try:
        with torch.no_grad():
            got = ops.conv_bias_act(x, w, b, 1, pad, dil, 1, act)
    finally:
        ops.set_timer(None)
    names = {r[0] for r in timer.records}
    assert names & {"conv3x3", "gemm_bf16"}, names          # the own kernel ran, not the library convolution
    assert got.shape == want.shape and got.dtype == torch.bfloat16
    _close(got, want, 6e-3, "frozen conv")


@pytest.mark.parametrize("shape", [(2, 64, 32, 48), (1, 128, 17, 31), (4, 8, 2, 2)])
def test_maxpool2x2_channels_last(shape):
    B, C, H, W = shape
    x = torch.randn(B, C, H, W, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        got = ops.max_pool2x2(x)
    assert torch.equal(got, F.max_pool2d(x, 2, 2))


@pytest.mark.parametrize("shape", [(8192, 320, 1280), (8192, 1280, 320), (2048, 512, 2048), (131072, 64, 64), (1000, 136, 72),
                                   (4096, 640, 320), (520, 2048, 512)])
def test_wgrad_gemm_with_fused_bias_gradient(shape):
    """dW += dy^T x with the bias gradient db += sum_t dy[t, :] reduced by the tensor core inside the same kernel (an extra
    N = 16 MMA per k-step against a tile of ones, 16 spare accumulator columns per buffer): both outputs accumulate on
    top of existing contents; ragged output-feature counts, one / many m-tiles, split contraction."""
    T, K_in, N_out = shape
    torch.manual_seed(T + K_in + N_out)
    x = torch.randn(T, K_in, device=DEV).bfloat16()
    dy = torch.randn(T, N_out, device=DEV).bfloat16()
    dw0 = torch.randn(N_out, K_in, device=DEV)
    db0 = torch.randn(N_out, device=DEV)
    dw, db = dw0.clone(), db0.clone()
    ops.gemm_bf16(dy, x, out=dw, a_mn_major=True, b_mn_major=True, accumulate=True, colsum_out=db)
    want_w = dw0 + dy.float().t() @ x.float()
    want_b = db0 + dy.float().sum(0)
    _close(dw, want_w, 2e-4, "dW")
    _close(db, want_b, 2e-5, "db")
