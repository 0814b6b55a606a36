#!/bin/bash
OUT=gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > $OUT/s11_pytest.log 2>&1; tail -3 $OUT/s11_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $OUT/s11_bench.json 2> $OUT/s11_bench.err
tail -3 $OUT/s11_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/s11_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['loss_src'])
P
timeout 300 python tools/profile_step.py --out $OUT/s11_step_profile.json > $OUT/s11_profile.log 2>&1
head -45 $OUT/s11_profile.log | cut -c1-150
