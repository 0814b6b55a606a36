import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops
T, K, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (8192, 320, 1280)
x = torch.randn(T, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
b = torch.randn(N, device="cuda"); dy = torch.randn(T, N, device="cuda").bfloat16(); g = torch.zeros(N, K, device="cuda")
for _ in range(3):
    ops.gemm_bf16(x, w, b)
    ops.gemm_bf16(dy, w, b_mn_major=True)
    ops.gemm_bf16(dy, x, out=g, a_mn_major=True, b_mn_major=True, accumulate=True)
    torch.nn.functional.linear(x, w, b.bfloat16())
torch.cuda.synchronize()
