#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_dacs_gpu.py -x -q > $OUT/r2_23_dacs.log 2>&1; echo dacs rc=$?
grep -v "^$" $OUT/r2_23_dacs.log | tail -15 | cut -c1-400
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2_23_pytest.log 2>&1; echo pytest rc=$?
tail -3 $OUT/r2_23_pytest.log
timeout 120 python tools/time_dacs.py > $OUT/r2_23_time_dacs.log 2>&1; tail -4 $OUT/r2_23_time_dacs.log
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/r2_23_bench.json 2> $OUT/r2_23_bench.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open("$OUT/r2_23_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step") if k in d}, d.get("e2e"))
PY
