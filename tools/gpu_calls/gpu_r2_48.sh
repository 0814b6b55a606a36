#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2_48_pytest.log 2>&1; echo pytest rc=$?
grep -v "^$" $OUT/r2_48_pytest.log | tail -6 | cut -c1-400
for own in 1 0; do
RF_OWN_GEMM=$own timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-corr-sweep > $OUT/r2_48_bench_own$own.json 2> $OUT/r2_48_bench_own$own.err; echo bench own=$own rc=$?
python - <<PY
import json
d=json.loads(open("$OUT/r2_48_bench_own$own.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches") if k in d}, d.get("e2e",{}).get("ms_per_step"))
PY
tail -2 $OUT/r2_48_bench_own$own.err | cut -c1-300
done
RF_OWN_GEMM=1 timeout 300 python tools/profile_step.py --out $OUT/r2_48_step_profile.json > $OUT/r2_48_profile.log 2>&1; tail -2 $OUT/r2_48_profile.log
