#!/usr/bin/env python
"""Kernel-level breakdown of one train step with torch.profiler (device times, not under ncu).
    python tools/profile_step.py [--size 1024] [--out gpurun_out/step_profile.json]
Writes the top kernels by device time and the launch count / GPU-busy fraction of the step."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--model", default="mit_b5")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "step_profile.json"))
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--ops", action="store_true", help="also list the top aten ops by device time with shapes + source line")
    ap.add_argument("--find", default="", help="comma-separated kernel-name substrings: print which ops (with parents) launch them")
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    model = bench.build_model(args.model, args.precision, dev)
    model.setup_runtime()
    batch = bench.synth_batch(args.size, 2, 100, dev)
    for i in range(3):
        model.training_step(batch, i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.training_step(batch, 3)
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=args.ops or bool(args.find),
                 with_stack=args.ops) as prof:
        model.training_step(batch, 4)
        torch.cuda.synchronize()
    rows = []
    total = 0.0
    n = 0
    for e in prof.key_averages():
        dt = getattr(e, "self_device_time_total", 0.0)
        if e.device_type.name == "CUDA" and dt > 0:
            rows.append((e.key, e.count, dt / 1e3))
            total += dt / 1e3
            n += e.count
    rows.sort(key=lambda r: -r[2])
    out = {"size": args.size, "model": args.model, "precision": args.precision, "step_wall_ms": t_wall * 1e3,
           "step_host_issue_ms": t_host * 1e3, "device_busy_ms": total, "device_launches": n,
           "top": [{"kernel": k[:160], "calls": c, "ms": round(ms, 3), "share": round(ms / total, 4)}
                   for k, c, ms in rows[:args.top]]}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "top"}))
    for r in out["top"][:40]:
        print("%8.3f ms %5.1f%% x%-5d %s" % (r["ms"], 100 * r["share"], r["calls"], r["kernel"][:110]))


    if args.find:
        pats = [p for p in args.find.split(",") if p]
        agg = {}
        for e in prof.events():
            for k in getattr(e, "kernels", []) or []:
                if all(p in k.name for p in pats):
                    chain, q = [], e
                    while q is not None and len(chain) < 5:
                        chain.append(q.name[:40])
                        q = getattr(q, "cpu_parent", None)
                    key = (" <- ".join(chain), str(getattr(e, "input_shapes", ""))[:80])
                    a = agg.setdefault(key, [0, 0.0])
                    a[0] += 1
                    a[1] += k.duration / 1e3
        for key, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            print("FIND %8.3f ms x%-4d %s | %s" % (ms, n, key[0], key[1]))
    if args.ops:
        ev = [e for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6)
              if getattr(e, "self_device_time_total", 0) > 0 and e.device_type.name != "CUDA"]
        ev.sort(key=lambda e: -e.self_device_time_total)
        for e in ev[:140]:
            stack = [l for l in (e.stack or []) if "refign_b200" in l or "bench.py" in l]
            print("%8.3f ms x%-5d %-28s %s | %s" % (e.self_device_time_total / 1e3, e.count, e.key[:28],
                                                  str(e.input_shapes)[:90], (stack[0].split("/")[-1] if stack else "")[:70]))


if __name__ == "__main__":
    main()
