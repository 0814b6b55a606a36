#!/bin/bash
OUT=gpurun_out
for v in explicit explicit_cs_template explicit explicit_cs_template; do
cp refign_b200/_alt/$v.so refign_b200/librefign_b200.so
echo "== $v"
timeout 300 python tools/bench_gemm.py 2>/dev/null | grep graph_replay
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-corr-sweep --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['avg_launch_us'])"
done
