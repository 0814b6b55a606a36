#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2_28_pytest.log 2>&1; echo pytest rc=$?
tail -2 $OUT/r2_28_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/r2_28_bench.json 2> $OUT/r2_28_bench.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open("$OUT/r2_28_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","roofline") if k in d}, d.get("e2e"))
PY
timeout 300 python tools/profile_step.py --out $OUT/r2_28_step_profile.json > $OUT/r2_28_profile.log 2>&1; tail -3 $OUT/r2_28_profile.log; ls $OUT | grep -i "step_profile" | tail -2
