"""Summaries of the ncu evidence under gpurun_out/ -> profiles/ (tracked):
  * the launch list (ncu --metrics gpu__time_duration.sum --csv) aggregated per kernel: launches, total ms, share
  * the `--set full` captures: a fixed set of raw metrics per captured kernel (ncu -i ... --page raw --csv).
    python tools/summarize_ncu.py <launches.csv> <out.csv> <out.json> <rep1.ncu-rep> [...]"""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(path, out, note):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.reader(io.StringIO("".join(rows)))
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(r[ki][:140], [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(v[1] for v in agg.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "share"])
        w.writerow(["# " + note, "", "", ""])
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, round(ms, 3), round(ms / tot, 4)])
    print("launch list: %d kernels, %d launches, %.1f ms" % (len(agg), sum(v[0] for v in agg.values()), tot))


def full(reps, out):
    res = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {"report": rep.split("/")[-1].replace(".ncu-rep", ""), "kernel": r[hdr.index("Kernel Name")][:110]}
            for m in WANT:
                if m in hdr:
                    i = hdr.index(m)
                    d[m] = "%s %s" % (r[i], units[i])
            res.append(d)
    json.dump(res, open(out, "w"), indent=0)
    print("full captures: %d kernels" % len(res))


if __name__ == "__main__":
    import os
    launches(sys.argv[1], sys.argv[2], os.environ.get("NCU_NOTE", "ncu --metrics gpu__time_duration.sum --clock-control none "
                                                      "<bench command>; cold-cache serialised times: compare SHARES"))
    full(sys.argv[4:], sys.argv[3])
