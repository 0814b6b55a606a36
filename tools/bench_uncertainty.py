"""Fused UncertaintyModule patch CNN at the alignment head's level shapes (B2 1024^2 input): kernel time via CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops
from refign_b200.modules import UncertaintyModule

for (s, H) in [(9, 256), (9, 128), (9, 32), (16, 16)]:
    m = UncertaintyModule(in_channels=1, search_size=s).cuda().eval()
    for p in m.parameters():
        p.requires_grad_(False)
    c = torch.rand(2, s * s, H, H, device="cuda")
    prm = m._fused_params()
    with torch.no_grad():
        for _ in range(3):
            ops.uncertainty_patch_cnn(c, prm, s, 0.1)
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.uncertainty_patch_cnn(c, prm, s, 0.1); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    print("search %2d  2x%dx%d: %.1f us" % (s, H, H, sorted(ts)[5] * 1e3), flush=True)
