"""Quick numerical check of the CTA-pair (cta_group::2) GEMM variants against fp32 matmul (RF_GEMM_PAIRS=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops
torch.manual_seed(0)
for (T, K, N) in [(256, 64, 128), (512, 128, 256), (8192, 320, 1280), (8192, 1280, 320), (1000, 72, 136), (2048, 512, 2048)]:
    x = torch.randn(T, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda"); dy = torch.randn(T, N, device="cuda").bfloat16()
    y = ops.gemm_bf16(x, w, b); torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + b
    e1 = float((y.float() - ref).abs().max() / ref.abs().max())
    dx = ops.gemm_bf16(dy, w, b_mn_major=True); torch.cuda.synchronize()
    ref = dy.float() @ w.float()
    e2 = float((dx.float() - ref).abs().max() / ref.abs().max())
    g = torch.zeros(N, K, device="cuda")
    if T % 8 == 0:
        ops.gemm_bf16(dy, x, out=g, a_mn_major=True, b_mn_major=True, accumulate=True); torch.cuda.synchronize()
        ref = dy.float().t() @ x.float()
        e3 = float((g - ref).abs().max() / ref.abs().max())
    else:
        e3 = -1
    print((T, K, N), "fwd %.2e dgrad %.2e wgrad %.2e" % (e1, e2, e3), flush=True)
