"""The BASELINE.json configurations that are not bench.py's default line, one JSON line each (CUDA events, synthetic
inputs, warm-up + timed iterations, inputs resident):
  config 2  UAWarpC alignment-only: AlignmentModel.forward on 8 pairs of 512x512 (VGG pyramids, global + local
            correlation volumes, flow + uncertainty decode) -- pairs/s
  config 3  DAFormer MiT-B5 + Refign train step at the 512x512 crop, 2 pairs + 2 source images per GPU -- through
            bench.py itself (``bench.py --size 512``; printed here for the record at N = 1)
  config 4  HRDA MiT-B5 multi-resolution + Refign train step at 1024x1024 (tools/bench_hrda.py, eager)
    python tools/bench_configs.py [--only 2,3,4]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def config2(precision="bf16", pairs=8, size=512, iters=10, warmup=3):
    import torch
    import refign_b200 as P
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = P.AlignmentModel(alignment_backbone=P.VGG('vgg16', out_indices=[2, 3, 4]),
                         alignment_head=P.UAWarpCHead(in_index=[0, 1], input_transform='multiple_select',
                                                      estimate_uncertainty=True),
                         precision=precision).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(3)
    i = torch.randn(pairs, 3, size, size, device=dev, generator=g)
    j = i.roll((3, -4), (2, 3)) + 0.05 * torch.randn(pairs, 3, size, size, device=dev, generator=g)
    with torch.no_grad():
        for _ in range(warmup):
            m(i, j)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            flow, unc = m(i, j)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"metric": "uawarpc_alignment_forward_image_pairs_per_s", "value": pairs / (ms * 1e-3), "unit": "pairs/s",
            "ms_per_step": ms, "steps": iters, "warmup": warmup, "n_gpus": 1, "dtype": precision, "data": "synthetic",
            "config": {"workload": "BASELINE config 2: AlignmentModel.forward, %d pairs of %dx%d" % (pairs, size, size)},
            "flow_shape": list(flow.shape), "finite": bool(torch.isfinite(flow).all() and torch.isfinite(unc).all())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="2,3,4")
    args = ap.parse_args()
    which = set(args.only.split(","))
    if "2" in which:
        print(json.dumps(config2()), flush=True)
    if "3" in which:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--size", "512", "--steps", "10", "--warmup", "3",
                              "--no-cpu-baseline", "--no-corr-sweep"], capture_output=True, text=True, cwd=ROOT)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if line:
            d = json.loads(line[-1])
            d.pop("own_kernels", None)
            d["config"]["workload"] = "BASELINE config 3: " + d["config"]["workload"]
            print(json.dumps(d), flush=True)
        else:
            print(json.dumps({"config": "3", "error": out.stderr[-400:]}), flush=True)
    if "4" in which:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_hrda.py"), "--steps", "3", "--warmup", "2"],
                             capture_output=True, text=True, cwd=ROOT)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        print(line[-1] if line else json.dumps({"config": "4", "error": out.stderr[-400:]}), flush=True)


if __name__ == "__main__":
    main()
