import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops
B, H, W, Ci, Co = (int(v) for v in sys.argv[1:6]) if len(sys.argv) > 5 else (2, 256, 256, 1024, 256)
x = torch.randn(B, H, W, Ci, device="cuda").bfloat16()
w = (torch.randn(Co, 3, 3, Ci, device="cuda") / (3 * Ci ** 0.5)).bfloat16()
dy = torch.randn(B, H, W, Co, device="cuda").bfloat16()
dw = torch.zeros(Co, 3, 3, Ci, device="cuda")
from refign_b200 import _lib
for _ in range(3):
    ops.conv3x3_nhwc_raw(x, w)
    _lib.check(_lib.lib().rf_conv3x3_wgrad_bf16(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), B, H, W, Ci, Co, torch.cuda.current_stream().cuda_stream), "wgrad")
torch.cuda.synchronize()
