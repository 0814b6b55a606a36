"""One forward + backward of the fused up-sampling / cross-entropy at the train-step shape (ncu capture target)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops
torch.manual_seed(0)
B, H = 2, 1024
lg = (3 * torch.randn(B, 19, H // 4, H // 4, device="cuda")).requires_grad_(True)
tg = torch.randint(0, 19, (B, H, H), device="cuda")
pw = torch.rand(B, H, H, device="cuda")
x4 = torch.randn(2 * B, 19, H // 4, H // 4, device="cuda")
for _ in range(2):
    loss = ops.upsample_cross_entropy(lg, tg, pw)
    loss.backward()
    ops.upsample_bilinear(x4, (H, H))
torch.cuda.synchronize()
