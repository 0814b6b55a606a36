"""One forward + a few backward launches of the fused SR-attention at a MiT-B5 stage shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops
stage = int(sys.argv[1]) if len(sys.argv) > 1 else 3
size = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
div, heads = ((4, 1), (8, 2), (16, 5), (32, 8))[stage - 1]
B, N, M = 2, (size // div) ** 2, (size // 32) ** 2
C = heads * 64
q = torch.randn(B, N, C, device="cuda").bfloat16().requires_grad_(True)
kv = torch.randn(B, M, 2 * C, device="cuda").bfloat16().requires_grad_(True)
o = ops.sr_attention(q, kv, heads, 0.125)
go = torch.randn_like(o)
for _ in range(4):
    torch.autograd.grad(o, (q, kv), go, retain_graph=True)
torch.cuda.synchronize()
