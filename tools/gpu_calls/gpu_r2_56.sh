#!/bin/bash
OUT=gpurun_out
timeout 120 tools/_bin/trace_gemm 8192 1280 320 > $OUT/r2_56_trace_a.txt; echo trace rc=$?
head -4 $OUT/r2_56_trace_a.txt | cut -c1-700
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q > $OUT/r2_56_gemm.log 2>&1; echo gemm rc=$?
grep -v "^$" $OUT/r2_56_gemm.log | tail -5 | cut -c1-300
timeout 300 python tools/bench_gemm.py > $OUT/r2_56_bench_gemm.jsonl 2> $OUT/r2_56_bench_gemm.err; echo rc=$?
grep "total" $OUT/r2_56_bench_gemm.jsonl | cut -c1-520; tail -2 $OUT/r2_56_bench_gemm.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-corr-sweep > $OUT/r2_56_bench.json 2> $OUT/r2_56_bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_56_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['roofline'])
for k,v in list(d['own_kernels'].items())[:10]: print(k, v)
PY
