#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/r2_14_pytest.log 2>&1; echo pytest rc=$?
tail -4 $OUT/r2_14_pytest.log
for st in 1 3; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sr_attention_bwd --csv --log-file $OUT/r2_14_launch_s$st.csv python tools/run_attn_bwd_once.py $st > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/r2_14_launch_s$st.csv")) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
for r in rows[-2:]: print("stage $st", r[ik][:40], r[iv])
PY
done
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/r2_14_bench.json 2> $OUT/r2_14_bench.err; echo bench rc=$?
python - <<PY
import json
d=json.loads(open("$OUT/r2_14_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","roofline") if k in d})
print(d.get("e2e"))
PY
