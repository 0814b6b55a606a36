#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q > $OUT/r2_54_gemm.log 2>&1; echo gemm rc=$?
grep -v "^$" $OUT/r2_54_gemm.log | tail -5 | cut -c1-300
timeout 300 python tools/bench_gemm.py > $OUT/r2_54_bench_gemm.jsonl 2> $OUT/r2_54_bench_gemm.err; echo rc=$?
grep "total" $OUT/r2_54_bench_gemm.jsonl | cut -c1-520; tail -2 $OUT/r2_54_bench_gemm.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-corr-sweep > $OUT/r2_54_bench.json 2> $OUT/r2_54_bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_54_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['roofline'])
for k,v in list(d['own_kernels'].items())[:8]: print(k, v)
PY
