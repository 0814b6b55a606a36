"""Fused up-sampling + cross-entropy backward at the train step's shape (B2, 19 x 256^2 -> 1024^2): tiled scatter kernel vs
the gather kernel (RF_UPSAMPLE_CE_BWD=gather), kernel time by CUDA events; the two gradients are compared."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for mode in ("tile", "gather"):
        env = dict(os.environ, RF_UPSAMPLE_CE_BWD=mode)
        subprocess.run([sys.executable, __file__, mode], env=env, check=True)
    import torch
    a, b = torch.load("/tmp/ce_grad_tile.pt"), torch.load("/tmp/ce_grad_gather.pt")
    print("max |tile - gather| = %.3e of max |grad| %.3e" % (float((a - b).abs().max()), float(b.abs().max())))
    sys.exit(0)
import torch
from refign_b200 import ops
torch.manual_seed(0)
lg = (torch.randn(2, 19, 256, 256, device="cuda") * 3).requires_grad_(True)
tg = torch.randint(0, 19, (2, 1024, 1024), device="cuda")
tg[:, :16] = 255
pw = torch.rand(2, 1024, 1024, device="cuda")
for _ in range(3):
    lg.grad = None
    ops.upsample_cross_entropy(lg, tg, pw, 255).backward()
timer = ops.KernelTimer()
ops.set_timer(timer)
for _ in range(10):
    lg.grad = None
    ops.upsample_cross_entropy(lg, tg, pw, 255).backward()
ops.set_timer(None)
torch.cuda.synchronize()
t = {}
for name, e0, e1, _ in timer.records:
    t.setdefault(name, []).append(e0.elapsed_time(e1) * 1e3)
print(sys.argv[1], {k: round(sorted(v)[len(v) // 2], 1) for k, v in t.items()}, "us per C-ABI call (event pair, ~5 us overhead)", flush=True)
torch.save(lg.grad.cpu(), "/tmp/ce_grad_%s.pt" % sys.argv[1])
