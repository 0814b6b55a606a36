"""add_layer_norm forward / backward kernels at the MiT-B5 stage shapes (B2 1024^2): device time of the C-ABI launches
(ops.KernelTimer event pairs, launch queue pre-filled so the host is never the bottleneck, L2 flushed per iteration)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops

flush = torch.empty(64 << 20, device="cuda", dtype=torch.float32)
over = ops.KernelTimer.calibrate(torch.device("cuda", 0))
for (N, C) in [(65536, 64), (16384, 128), (4096, 320), (1024, 512)]:
    ln = torch.nn.LayerNorm(C).cuda()
    x = torch.randn(2, N, C, device="cuda", requires_grad=True)
    br = torch.randn(2, N, C, device="cuda").bfloat16().requires_grad_(True)
    gx, gy = torch.randn(2, N, C, device="cuda"), torch.randn(2, N, C, device="cuda").bfloat16()
    for _ in range(3):
        xn, y = ops.add_layer_norm(x, br, None, ln, out_dtype=torch.bfloat16)
        torch.autograd.backward((xn, y), (gx, gy))
    timer = ops.KernelTimer()
    ops.set_timer(timer)
    torch.cuda._sleep(int(0.05 * 1.9e9))
    for _ in range(20):
        flush.zero_()
        xn, y = ops.add_layer_norm(x, br, None, ln, out_dtype=torch.bfloat16)
        torch.autograd.backward((xn, y), (gx, gy))
    ops.set_timer(None)
    torch.cuda.synchronize()
    t = {}
    for name, e0, e1, _ in timer.records:
        t.setdefault(name, []).append(e0.elapsed_time(e1) - over)
    rows = 2 * N
    med = {k: sorted(v)[len(v) // 2] * 1e3 for k, v in t.items()}
    print("rows %6d C %3d: " % (rows, C) + "  ".join("%s %.1f us (%.0f GB/s)" % (k, v, rows * C * (14 if 'fwd' in k else 16) / v / 1e3)
                                                  for k, v in med.items()), flush=True)
