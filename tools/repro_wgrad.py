import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops
torch.manual_seed(0)
for (B, H, W, C, dil, dt) in [(2, 128, 128, 1024, 6, torch.bfloat16), (2, 128, 128, 1024, 18, torch.bfloat16), (2, 128, 128, 128, 1, torch.bfloat16), (2, 64, 64, 256, 1, torch.bfloat16), (2, 32, 32, 640, 1, torch.bfloat16), (2,16,16,1024,1,torch.bfloat16)]:
    x = torch.randn(B, H, W, C, device="cuda").to(dt).requires_grad_(True)
    w = torch.randn(C, 1, 3, 3, device="cuda", requires_grad=True)
    b = torch.randn(C, device="cuda", requires_grad=True) if dil == 1 else None
    if dil == 1:
        y = ops.dwconv3x3_gelu(x.view(B, H * W, C), H, W, w, b)
    else:
        y = ops.dwconv3x3_nhwc(x.permute(0, 3, 1, 2), w, None, dil)
    y.float().sum().backward()
    torch.cuda.synchronize()
    print("ok", B, H, W, C, dil, float(w.grad.abs().sum()))
