#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2_83_pytest.log 2>&1; echo pytest rc=$?
tail -3 $OUT/r2_83_pytest.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $OUT/r2_83_bench.json 2> $OUT/r2_83_bench.err; echo bench rc=$?
tail -2 $OUT/r2_83_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_83_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'])
print(d['roofline'])
print(d['corr_volume_GBps'])
print({k:v for k,v in d['cpu_baseline'].items() if k!='corr_volume'})
print(d['clocks'])
PY
timeout 300 python tools/profile_step.py --out $OUT/r2_83_step_profile.json > $OUT/r2_83_profile.log 2>&1; tail -1 $OUT/r2_83_profile.log
