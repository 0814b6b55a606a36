#!/bin/bash
OUT=gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-corr-sweep > $OUT/r2_59_bench.json 2> $OUT/r2_59_bench.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_59_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'] and d['e2e']['value'], d['roofline'])
for k,v in list(d['own_kernels'].items())[:12]: print(k, v)
PY
timeout 300 python tools/profile_step.py --out $OUT/r2_59_step_profile.json > $OUT/r2_59_profile.log 2>&1; tail -2 $OUT/r2_59_profile.log
RF_PROFILER_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-corr-sweep --no-e2e > $OUT/r02_launches_bench.log 2>&1; echo launches rc=$?
wc -l $OUT/r02_launches_bench.csv
