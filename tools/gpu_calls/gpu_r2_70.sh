#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_hrda_gpu.py -x -q > $OUT/r2_70_hrda.log 2>&1; echo hrda rc=$?
tail -25 $OUT/r2_70_hrda.log | cut -c1-400
timeout 600 python bench.py --workload hrda --steps 5 --warmup 3 --no-corr-sweep > $OUT/r2_70_bench_hrda.json 2> $OUT/r2_70_bench_hrda.err; echo bench rc=$?
tail -3 $OUT/r2_70_bench_hrda.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_70_bench_hrda.json').read().strip().splitlines()[-1])
    print(d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['config'])
except Exception as e: print('no json', e)
PY
