import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops
B, N, M, heads = 2, 65536, 1024, 1
C = heads * 64
q = torch.randn(B, N, C, device="cuda").bfloat16()
kv = torch.randn(B, M, 2 * C, device="cuda").bfloat16()
for _ in range(3):
    ops.sr_attention_fwd(q, kv, heads, 0.125, want_lse=True)
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "bwd":
    qg = q.clone().requires_grad_(True); kvg = kv.clone().requires_grad_(True)
    o = ops.sr_attention(qg, kvg, heads, 0.125)
    for _ in range(3):
        torch.autograd.grad(o, (qg, kvg), torch.randn_like(o), retain_graph=True)
    torch.cuda.synchronize()
