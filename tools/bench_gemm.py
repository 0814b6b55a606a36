"""Own tcgen05 GEMM vs the library GEMM at the Linear-layer shapes of one MiT-B5 train step at 1024x1024
(CUDA events, L2 flushed, median of 10): forward (bias), dgrad, wgrad (fp32 accumulate)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def med(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

def graphed(fn, reps=10, iters=5):
    """device time per call when replayed from a CUDA graph (no host launch / tensor-map encoding cost), back to back"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    ts.sort()
    return ts[len(ts) // 2]


# (tokens, in, out): q/proj, kv (after SR), fc1, fc2 per stage at 1024^2, B = 2
stages = [(2 * 65536, 64), (2 * 16384, 128), (2 * 4096, 320), (2 * 1024, 512)]
shapes = []
for T, d in stages:
    shapes += [(T, d, d), (2 * 1024, d, 2 * d), (T, d, 4 * d), (T, 4 * d, d)]
tot = {"own": [0, 0, 0], "lib": [0, 0, 0]}
gt = {"own": [0, 0, 0], "lib": [0, 0, 0]}
for T, K, N in shapes:
    x = torch.randn(T, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev); bb = b.bfloat16(); dy = torch.randn(T, N, device=dev).bfloat16()
    g = torch.zeros(N, K, device=dev)
    r = {"T": T, "K": K, "N": N,
         "fwd_own": med(lambda: ops.gemm_bf16(x, w, b)), "fwd_lib": med(lambda: torch.nn.functional.linear(x, w, bb)),
         "dgrad_own": med(lambda: ops.gemm_bf16(dy, w, b_mn_major=True)), "dgrad_lib": med(lambda: dy @ w),
         "wgrad_own": med(lambda: ops.gemm_bf16(dy, x, out=g, a_mn_major=True, b_mn_major=True, accumulate=True)),
         "wgrad_lib": med(lambda: ops._mm_f32_acc(g, dy.t(), x))}
    r["g_fwd_own"] = graphed(lambda: ops.gemm_bf16(x, w, b)); r["g_fwd_lib"] = graphed(lambda: torch.nn.functional.linear(x, w, bb))
    r["g_dgrad_own"] = graphed(lambda: ops.gemm_bf16(dy, w, b_mn_major=True)); r["g_dgrad_lib"] = graphed(lambda: dy @ w)
    r["g_wgrad_own"] = graphed(lambda: ops.gemm_bf16(dy, x, out=g, a_mn_major=True, b_mn_major=True, accumulate=True))
    r["g_wgrad_lib"] = graphed(lambda: ops._mm_f32_acc(g, dy.t(), x))
    for i, k in enumerate(("g_fwd", "g_dgrad", "g_wgrad")):
        gt["own"][i] += r[k + "_own"]; gt["lib"][i] += r[k + "_lib"]
    fl = 2.0 * T * K * N
    r["fwd_own_TFLOPs"] = round(fl / r["fwd_own"] / 1e6, 1)
    for i, k in enumerate(("fwd", "dgrad", "wgrad")):
        tot["own"][i] += r[k + "_own"]; tot["lib"][i] += r[k + "_lib"]
    print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()}))
print(json.dumps({"graph_replay_total_us_own_fwd_dgrad_wgrad": [round(v, 1) for v in gt["own"]], "graph_replay_total_us_lib": [round(v, 1) for v in gt["lib"]]}))
print(json.dumps({"total_us_own_fwd_dgrad_wgrad": [round(v, 1) for v in tot["own"]], "total_us_lib": [round(v, 1) for v in tot["lib"]]}))
