"""Global-correlation sweep-size kernel in isolation: per-mode timing (mode bit 0 = mutual matching -> phase 0,
bit 1 = L2-norm -> phase 1; phase 2 always runs) or a single call for an ncu capture.
    python tools/run_gcorr_once.py time [N] [tc]      python tools/run_gcorr_once.py once [N] [tc]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops

what = sys.argv[1] if len(sys.argv) > 1 else "time"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
tc = int(sys.argv[3]) if len(sys.argv) > 3 else 2
C = int(sys.argv[4]) if len(sys.argv) > 4 else 128
unit = lambda x: torch.nn.functional.normalize(x, p=2, dim=1)
torch.manual_seed(0)
s, t = unit(torch.randn(1, C, N, N, device="cuda")), unit(torch.randn(1, C, N, N, device="cuda"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(mm, nrm):
    return ops.global_correlation(s, t, cyclic_consistency=mm, normalise=nrm, use_tensor_cores=tc)


if what == "once":
    run(True, True)
    torch.cuda.synchronize()
    run(True, True)
    torch.cuda.synchronize()
else:
    for mm, nrm in [(False, False), (True, False), (False, True), (True, True)]:
        for _ in range(3):
            run(mm, nrm)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(mm, nrm)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print("N=%d C=%d tc=%d mm=%d nrm=%d: median %.1f us (min %.1f)" % (N, C, tc, mm, nrm, ts[5], ts[0]), flush=True)
