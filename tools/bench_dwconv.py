"""Depthwise-conv / colsum kernel timings at the MiT-B5 1024x1024 Mix-FFN shapes (B=2 per pass).

    python tools/bench_dwconv.py [--iters 20] [--hot]   # --hot: no L2 flush (operands L2-resident, as inside the step)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from refign_b200 import ops  # noqa: E402
from refign_b200._lib import ptr  # noqa: E402

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda", dtype=torch.float32)
    _flush.zero_()

SHAPES = [(2, 256, 256, 256), (2, 128, 128, 512), (2, 64, 64, 1280), (2, 32, 32, 2048)]  # B, H, W, hidden C


def timeit(fn, iters, hot):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if not hot:
            flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--hot", action="store_true")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--only", default="")
    ap.add_argument("--aspp", action="store_true")
    args = ap.parse_args()
    if args.aspp:
        return aspp(args)
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    L = ops._lib.lib()
    for (B, H, W, C) in SHAPES:
        if args.only and str(C) != args.only:
            continue
        x = torch.randn(B, H, W, C, device="cuda", dtype=dt)
        gy = torch.randn_like(x)
        w = torch.randn(C, 1, 3, 3, device="cuda") * 0.2
        b = torch.randn(C, device="cuda") * 0.1
        y, g, gx = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        gw, gb = torch.empty_like(w), torch.empty_like(b)
        code = 1 if dt == torch.bfloat16 else 0
        st = ops._stream()
        n = x.numel() * x.element_size()
        P = ptr
        runs = {
            "fwd_gelu": (lambda: L.rf_dwconv3x3_nhwc_fwd(P(x), P(w), P(b), P(y), B, H, W, C, 1, 1, code, st), 2 * n),
            "gelu_bwd_pre": (lambda: L.rf_dwconv3x3_gelu_bwd_pre(P(x), P(w), P(b), P(gy), P(g), B, H, W, C, 1, code, st), 3 * n),
            "bwd_input": (lambda: L.rf_dwconv3x3_nhwc_bwd_input(P(gy), P(w), P(gx), B, H, W, C, 1, code, st), 2 * n),
            "bwd_weight": (lambda: L.rf_dwconv3x3_nhwc_bwd_weight(P(x), P(gy), P(gw), P(gb), B, H, W, C, 1, code, 0, st), 2 * n),
        }
        x2 = x.view(-1, C)
        cs = torch.empty(C, device="cuda")
        runs["colsum"] = (lambda: L.rf_colsum(P(x2), P(cs), x2.shape[0], C, code, 0, st), n)
        for name, (fn, nbytes) in runs.items():
            sec = timeit(fn, args.iters, args.hot)
            print(json.dumps({"kernel": name, "shape": [B, H, W, C], "dtype": args.dtype, "hot_l2": args.hot,
                              "us": round(sec * 1e6, 2), "alg_MB": round(nbytes / 1e6, 2),
                              "GBps": round(nbytes / sec / 1e9, 1)}), flush=True)


def aspp(args):
    """DAFormer ASPP depthwise branches at 1024x1024: [B,256,256,1024] bf16, dilation 6 / 12 / 18."""
    L = ops._lib.lib()
    B, H, W, C = 2, 256, 256, 1024
    x = torch.randn(B, H, W, C, device="cuda", dtype=torch.bfloat16)
    gy = torch.randn_like(x)
    y = torch.empty_like(x)
    w = torch.randn(C, 1, 3, 3, device="cuda") * 0.2
    gw = torch.empty_like(w)
    st = ops._stream()
    n = x.numel() * 2
    P = ptr
    for d in (6, 12, 18):
        runs = {"fwd": lambda: L.rf_dwconv3x3_nhwc_fwd(P(x), P(w), None, P(y), B, H, W, C, d, 0, 1, st),
                "bwd_input": lambda: L.rf_dwconv3x3_nhwc_bwd_input(P(gy), P(w), P(y), B, H, W, C, d, 1, st),
                "bwd_weight": lambda: L.rf_dwconv3x3_nhwc_bwd_weight(P(x), P(gy), P(gw), None, B, H, W, C, d, 1, 0, st)}
        for name, fn in runs.items():
            sec = timeit(fn, args.iters, args.hot)
            print(json.dumps({"kernel": "aspp_" + name, "dil": d, "shape": [B, H, W, C], "us": round(sec * 1e6, 2),
                              "alg_MB": round(2 * n / 1e6, 1), "GBps": round(2 * n / sec / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
