#!/bin/bash
OUT=gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > $OUT/s10_pytest.log 2>&1; tail -3 $OUT/s10_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $OUT/s10_bench.json 2> $OUT/s10_bench.err
tail -3 $OUT/s10_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/s10_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['loss_src'])
P
