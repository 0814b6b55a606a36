import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from refign_b200 import ops
from refign_b200.modules import UncertaintyModule
a = F.normalize(torch.randn(2, 128, 256, 256, device="cuda"), dim=1)
b = F.normalize(torch.randn(2, 128, 256, 256, device="cuda"), dim=1)
m = UncertaintyModule(in_channels=1, search_size=9, feed_in_previous=False).cuda().eval()
for p in m.parameters():
    p.requires_grad_(False)
x = torch.randn(8192, 320, device="cuda").bfloat16(); dy = torch.randn(8192, 1280, device="cuda").bfloat16()
dw = torch.zeros(1280, 320, device="cuda"); db = torch.zeros(1280, device="cuda")
with torch.no_grad():
    for _ in range(3):
        c = ops.local_correlation_relu_l2norm(a, b, 9)
        ops.uncertainty_patch_cnn(c, m._fused_params(), 9, 0.1)
        ops.gemm_bf16(dy, x, out=dw, a_mn_major=True, b_mn_major=True, accumulate=True, colsum_out=db)
torch.cuda.synchronize()
