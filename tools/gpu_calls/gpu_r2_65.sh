#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2_65_pytest.log 2>&1; echo pytest rc=$?
tail -3 $OUT/r2_65_pytest.log | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-corr-sweep > $OUT/r2_65_bench.json 2> $OUT/r2_65_bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_65_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['gpu_launches'], d['roofline'])
for k,v in list(d['own_kernels'].items())[:14]: print(k, v)
PY
