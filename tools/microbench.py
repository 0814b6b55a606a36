"""Per-kernel timing on one B200: CUDA events on the launching stream, L2 flushed between
iterations, median of N.  Prints one JSON line per kernel with the algorithmic bytes/flops
(DESIGN.md formulas) and the achieved fraction of the measured peaks.

    python tools/microbench.py [--iters 30] [--out gpurun_out/microbench.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from refign_b200 import ops  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda", dtype=torch.float32)
    _flush.zero_()


def time_op(fn, iters=30, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def unit(x):
    return torch.nn.functional.normalize(x, p=2, dim=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    hbm, tfl, src = peaks()
    dev = "cuda"
    res = []

    def report(name, shape, sec, nbytes, flops, bound="hbm"):
        r = {"kernel": name, "shape": shape, "us": round(sec * 1e6, 2), "alg_GB": round(nbytes / 1e9, 4),
             "GBps": round(nbytes / sec / 1e9, 1), "hbm_frac": round(nbytes / sec / 1e9 / hbm, 4),
             "GFLOP": round(flops / 1e9, 3), "TFLOPs": round(flops / sec / 1e12, 2), "bound": bound, "peaks": src}
        res.append(r)
        print(json.dumps(r), flush=True)

    torch.manual_seed(0)
    want = lambda k: (not args.only) or (args.only in k)
    if want("local_corr"):
        for (B, C, H, P) in [(1, 128, 64, 9), (2, 256, 32, 9), (2, 256, 64, 9), (2, 128, 128, 9), (2, 256, 128, 9),
                             (2, 128, 256, 9), (8, 128, 128, 9)]:
            a, b = unit(torch.randn(B, C, H, H, device=dev)), unit(torch.randn(B, C, H, H, device=dev))
            nbytes = 4 * B * H * H * (2 * C + P * P)
            flops = 2 * B * H * H * P * P * C
            report("local_corr_fwd+relu_l2norm", [B, C, H, H, P], time_op(lambda: ops.local_correlation_relu_l2norm(b, a, P), args.iters), nbytes, flops)
            report("local_corr_fwd", [B, C, H, H, P], time_op(lambda: ops.spatial_correlation_sample(a, b, patch_size=P), args.iters), nbytes, flops)
        for (B, C, H, P) in [(2, 128, 128, 9), (2, 128, 256, 9), (2, 256, 64, 9)]:
            a, b = unit(torch.randn(B, C, H, H, device=dev)).requires_grad_(True), unit(torch.randn(B, C, H, H, device=dev)).requires_grad_(True)
            out = ops.spatial_correlation_sample(a, b, patch_size=P)
            g = torch.randn_like(out)
            report("local_corr_bwd(both grads)", [B, C, H, H, P], time_op(lambda: torch.autograd.grad(out, (a, b), g, retain_graph=True), max(5, args.iters // 3)),
                   4 * B * H * H * (4 * C + P * P), 4 * B * H * H * P * P * C)
    if want("global_corr"):
        for (B, C, N) in [(2, 512, 16), (1, 128, 64), (1, 128, 128)]:
            s, t = unit(torch.randn(B, C, N, N, device=dev)), unit(torch.randn(B, C, N, N, device=dev))
            nn_ = N * N
            report("global_corr(ffma)+mm+relu_l2norm", [B, C, N, N], time_op(lambda: ops.global_correlation(s, t, use_tensor_cores=0), args.iters),
                   4 * B * (C * 2 * nn_ + nn_ * nn_), 2 * B * nn_ * nn_ * C)
    if want("global_tc"):
        for (B, C, N) in [(1, 128, 64), (1, 128, 128), (2, 128, 128), (1, 128, 192)]:
            s, t = unit(torch.randn(B, C, N, N, device=dev)), unit(torch.randn(B, C, N, N, device=dev))
            nn_ = N * N
            report("global_corr(tcgen05 tf32)+mm+relu_l2norm", [B, C, N, N], time_op(lambda: ops.global_correlation(s, t, use_tensor_cores=1), args.iters),
                   4 * B * (C * 2 * nn_ + nn_ * nn_), 3 * 2 * B * nn_ * nn_ * C)
    if want("global_pp"):
        for (B, C, N) in [(1, 128, 64), (1, 128, 128), (2, 128, 128), (1, 128, 192), (1, 64, 128)]:
            s, t = unit(torch.randn(B, C, N, N, device=dev)), unit(torch.randn(B, C, N, N, device=dev))
            nn_ = N * N
            report("global_corr(persistent tcgen05 tf32)+mm+relu_l2norm", [B, C, N, N], time_op(lambda: ops.global_correlation(s, t, use_tensor_cores=2), args.iters),
                   4 * B * (C * 2 * nn_ + nn_ * nn_), 4 * 2 * B * nn_ * nn_ * C)   # 4 tile passes: 2 x row max, row norm, write
            report("global_corr(persistent tcgen05 tf32) raw volume", [B, C, N, N], time_op(lambda: ops.global_correlation(s, t, cyclic_consistency=False, normalise=False, use_tensor_cores=2), args.iters),
                   4 * B * (C * 2 * nn_ + nn_ * nn_), 2 * B * nn_ * nn_ * C)
    if want("upsample_ce"):
        for (B, H) in [(2, 1024), (2, 512)]:
            lg = (3 * torch.randn(B, 19, H // 4, H // 4, device=dev)).requires_grad_(True)
            tg = torch.randint(0, 19, (B, H, H), device=dev)
            pw = torch.rand(B, H, H, device=dev)
            report("upsample_ce_fwd", [B, 19, H, H], time_op(lambda: ops.upsample_cross_entropy(lg.detach(), tg, pw), args.iters),
                   4 * lg.numel() + 12 * B * H * H, 12 * B * H * H * 19)
            loss = ops.upsample_cross_entropy(lg, tg, pw)
            report("upsample_ce_bwd", [B, 19, H, H], time_op(lambda: torch.autograd.grad(loss, lg, retain_graph=True), args.iters),
                   8 * lg.numel() + 12 * B * H * H, 4 * 24 * B * H * H * 19)
            x4 = torch.randn(2 * B, 19, H // 4, H // 4, device=dev)
            report("upsample_bilinear_f32", [2 * B, 19, H, H], time_op(lambda: ops.upsample_bilinear(x4, (H, H)), args.iters),
                   4 * x4.numel() * 17, 8 * x4.numel() * 16)
            up = lambda: torch.nn.functional.cross_entropy(torch.nn.functional.interpolate(lg.detach(), (H, H), mode="bilinear", align_corners=False), tg)
            report("library interpolate+cross_entropy fwd (for comparison)", [B, 19, H, H], time_op(up, args.iters), 4 * lg.numel() + 12 * B * H * H, 0)
    if want("warp"):
        for (B, C, H) in [(2, 19, 512), (2, 19, 1024), (2, 128, 256), (2, 256, 128)]:
            x, f = torch.randn(B, C, H, H, device=dev), torch.randn(B, 2, H, H, device=dev) * 4
            report("warp+mask(noise flow, sigma 4 px)", [B, C, H, H], time_op(lambda: ops.warp(x, f, return_mask=True), args.iters),
                   4 * B * H * H * (2 * C + 2) + B * H * H, 8 * B * C * H * H)
            # smooth flow (what the UAWarpC head produces: a 1/4-resolution field upsampled bilinearly)
            fs = torch.nn.functional.interpolate(torch.randn(B, 2, H // 32, H // 32, device=dev) * 8, size=(H, H),
                                                 mode="bilinear") + 3.0
            report("warp+mask(smooth flow)", [B, C, H, H], time_op(lambda: ops.warp(x, fs, return_mask=True), args.iters),
                   4 * B * H * H * (2 * C + 2) + B * H * H, 8 * B * C * H * H)
    if want("refine"):
        for H in [512, 1024]:
            lt, lr = torch.randn(2, 19, H, H, device=dev) * 3, torch.randn(2, 19, H, H, device=dev) * 3
            m, lv = torch.rand(2, H, H, device=dev) > 0.1, torch.randn(2, 1, H, H, device=dev)
            report("refine+cert+argmax", [2, 19, H, H], time_op(lambda: ops.refine_fused(lt, lr, m, logvar=lv), args.iters),
                   4 * 2 * H * H * (3 * 19 + 1) + 2 * H * H * (1 + 8 + 4), 0)
    if want("patch_embed"):
        import torch.nn as nn
        for (B, H) in [(2, 1024), (4, 1024), (2, 512)]:
            x = torch.randn(B, 3, H, H, device=dev)
            conv = nn.Conv2d(3, 64, 7, 4, 3).to(dev)
            ln = nn.LayerNorm(64).to(dev)
            with torch.no_grad():
                report("patch_embed(conv7x7s4+LN)", [B, 3, H, H], time_op(lambda: ops.patch_embed_ln(x, conv.weight, conv.bias, ln.weight, ln.bias, 1e-5), args.iters),
                       4 * B * (3 * H * H + 64 * H * H // 16), 2 * B * (H * H // 16) * 64 * 147)
    if want("optim"):
        n = 85_160_000
        p, g, m, v, e = (torch.randn(n, device=dev) for _ in range(5))
        v.abs_()
        report("ema", [n], time_op(lambda: ops.ema_update_(e, p, 0.999), args.iters), 12 * n, 3 * n)
        report("adamw", [n], time_op(lambda: ops.adamw_step_(p, g, m, v, [n // 20, n // 10, n - 1000, n], [6e-4, 6e-4, 6e-5, 6e-5],
                                                               [0.01, 0, 0.01, 0], 0.9, 0.999, 1e-8, 5), args.iters), 28 * n, 12 * n)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
