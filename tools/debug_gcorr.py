import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops
torch.manual_seed(0)
def unit(x): return torch.nn.functional.normalize(x, p=2, dim=1)
for (C, Hs, Ws, Ht, Wt) in [(32, 8, 16, 8, 16), (128, 8, 16, 8, 16), (128, 16, 16, 16, 16)]:
    s = unit(torch.randn(1, C, Hs, Ws, device="cuda")); t = unit(torch.randn(1, C, Ht, Wt, device="cuda"))
    ref = torch.einsum("bcs,bct->bst", s.flatten(2), t.flatten(2))
    got = ops.global_correlation(s, t, cyclic_consistency=False, normalise=False, use_tensor_cores=1).flatten(2)
    err = (got - ref).abs()
    print("C", C, "Ns", Hs * Ws, "Nt", Ht * Wt, "max err", float(err.max()), "max ref", float(ref.abs().max()), "got absmax", float(got.abs().max()))
    # error by 32x32 sub-block
    Ns, Nt = ref.shape[1], ref.shape[2]
    blk = err[0].reshape(Ns // 32, 32, Nt // 32, 32).amax(dim=(1, 3))
    print(blk)
    # does got equal ref under some permutation of 32-blocks? check got block (i,j) against ref block candidates
    g = got[0].reshape(Ns // 32, 32, Nt // 32, 32); r = ref[0].reshape(Ns // 32, 32, Nt // 32, 32)
    for i in range(min(2, Ns // 32)):
        for j in range(min(2, Nt // 32)):
            best = min(((float((g[i, :, j, :] - r[a, :, b_, :]).abs().max()), a, b_) for a in range(Ns // 32) for b_ in range(Nt // 32)))
            bestT = min(((float((g[i, :, j, :] - r[a, :, b_, :].t()).abs().max()), a, b_) for a in range(Ns // 32) for b_ in range(Nt // 32)))
            print("  got block", (i, j), "closest ref block", best, "closest transposed", bestT)
