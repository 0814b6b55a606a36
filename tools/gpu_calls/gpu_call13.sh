#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/s13_pytest.log 2>&1; tail -6 $OUT/s13_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $OUT/s13_bench.json 2> $OUT/s13_bench.err
tail -3 $OUT/s13_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/s13_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['loss_src'])
P
