#!/bin/bash
OUT=gpurun_out
RF_GEMM_PAIRS=1 timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q > $OUT/r2_60_gemm.log 2>&1; echo gemm rc=$?
grep -v "^$" $OUT/r2_60_gemm.log | tail -3 | cut -c1-300
RF_GEMM_PAIRS=1 timeout 300 python tools/bench_gemm.py > $OUT/r2_60_bench_gemm.jsonl 2> $OUT/r2_60_bench_gemm.err; echo rc=$?
grep "total" $OUT/r2_60_bench_gemm.jsonl | cut -c1-520; tail -2 $OUT/r2_60_bench_gemm.err
for pr in 1 0; do
RF_GEMM_PAIRS=$pr timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-corr-sweep > $OUT/r2_60_bench_pairs$pr.json 2> $OUT/r2_60_bench_pairs$pr.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_60_bench_pairs$pr.json').read().strip().splitlines()[-1])
print($pr, d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['roofline'])
for k,v in list(d['own_kernels'].items())[:6]: print(k, v)
PY
done
