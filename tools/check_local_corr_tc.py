"""Tensor-core local correlation (RF_LOCAL_CORR_TC, default on) against the exact-fp32 FFMA tiles: error and time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from refign_b200 import ops


def run(a, b, fused, tc):
    os.environ["RF_LOCAL_CORR_TC"] = "1" if tc else "0"
    if fused:
        return ops.local_correlation_relu_l2norm(a, b, 9)
    return ops.spatial_correlation_sample(a, b, patch_size=9)


def timeit(fn, n=20):
    flush = torch.empty(64 << 20, device="cuda", dtype=torch.float32)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


torch.manual_seed(0)
for (B, C, H, W) in [(2, 128, 256, 256), (1, 128, 64, 64), (2, 64, 100, 68), (1, 32, 67, 36), (2, 128, 128, 128)]:
    a = F.normalize(torch.randn(B, C, H, W, device="cuda"), dim=1)
    b = F.normalize(torch.randn(B, C, H, W, device="cuda"), dim=1)
    for fused in (False, True):
        ref = run(a, b, fused, False)
        got = run(a, b, fused, True)
        torch.cuda.synchronize()
        err = float((got - ref).abs().max())
        t0 = timeit(lambda: run(a, b, fused, False))
        t1 = timeit(lambda: run(a, b, fused, True))
        print((B, C, H, W), "fused" if fused else "plain", "max|err| %.2e (max|ref| %.2e)  ffma %.1f us  tc %.1f us" % (err, float(ref.abs().max()), t0, t1), flush=True)
