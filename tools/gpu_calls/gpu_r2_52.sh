#!/bin/bash
OUT=gpurun_out
for pairs in 1 0; do
RF_GEMM_PAIRS=$pairs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-corr-sweep > $OUT/r2_52_bench_pairs$pairs.json 2> $OUT/r2_52_bench_pairs$pairs.err; echo bench pairs=$pairs rc=$?
python - <<PY
import json
d=json.loads(open("$OUT/r2_52_bench_pairs$pairs.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches") if k in d}, d.get("e2e",{}).get("ms_per_step"))
PY
done
