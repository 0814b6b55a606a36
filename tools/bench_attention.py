"""Device timing of the fused SR-attention forward at the MiT-B5 shapes (CUDA events, L2 flushed)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refign_b200 import ops

def timeit(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 1590.0
FAST = "--fast" in sys.argv   # 1024 only, no library baselines
for size in ((1024,) if FAST else (512, 1024)):
    for stage, (div, heads) in enumerate(((4, 1), (8, 2), (16, 5), (32, 8))):
        B = 2
        N = (size // div) ** 2
        M = (size // 32) ** 2
        C = heads * 64
        q = torch.randn(B, N, C, device="cuda").bfloat16()
        kv = torch.randn(B, M, 2 * C, device="cuda").bfloat16()
        ms = timeit(lambda: ops.sr_attention_fwd(q, kv, heads, 0.125))
        ms_lib = 0.0 if FAST else timeit(lambda: ops._sr_attention_library(q, kv, heads, 0.125))
        qg = q.clone().requires_grad_(True); kvg = kv.clone().requires_grad_(True)
        o = ops.sr_attention(qg, kvg, heads, 0.125); go = torch.randn_like(o)
        ms_bwd = timeit(lambda: torch.autograd.grad(o, (qg, kvg), go, retain_graph=True))
        ms_bwd_lib = 0.0
        if not FAST:
            ol = ops._sr_attention_library(qg, kvg, heads, 0.125)
            ms_bwd_lib = timeit(lambda: torch.autograd.grad(ol, (qg, kvg), go, retain_graph=True))
        fl = 4.0 * B * heads * N * M * 64
        print(json.dumps({"kernel": "sr_attention_fwd", "size": size, "stage": stage + 1, "B": B, "N": N, "M": M, "heads": heads,
                          "us": round(ms * 1e3, 1), "TFLOPs": round(fl / ms / 1e9, 1), "frac_of_bf16_peak": round(fl / ms / 1e9 / peak, 4),
                          "library_us": round(ms_lib * 1e3, 1),
                          "bwd_us": round(ms_bwd * 1e3, 1), "bwd_TFLOPs_algorithmic_5gemm": round(2.5 * fl / ms_bwd / 1e9, 1), "bwd_frac_of_bf16_peak": round(2.5 * fl / ms_bwd / 1e9 / peak, 4), "bwd_library_us": round(ms_bwd_lib * 1e3, 1)}))
