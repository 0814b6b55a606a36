#!/bin/bash
OUT=gpurun_out
( time timeout 600 python bench.py --impl reference --cpu-size 1024 --steps 1 --warmup 1 > $OUT/r2_31_ref1024.json 2> $OUT/r2_31_ref1024.err ) 2> $OUT/r2_31_time.txt; echo rc=$?
cut -c1-700 $OUT/r2_31_ref1024.json; tail -3 $OUT/r2_31_time.txt; free -g | head -2; nproc
timeout 600 python bench.py --steps 6 --warmup 3 > $OUT/r2_31_bench.json 2> $OUT/r2_31_bench.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open("$OUT/r2_31_bench.json").read().strip().splitlines()[-1])
print(d["corr_volume_GBps"]); print(d["cpu_baseline"])
PY
