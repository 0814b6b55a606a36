#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_mit_ops_gpu.py -x -q -k "sr_attention" > $OUT/r2_27_pytest.log 2>&1; echo pytest rc=$?
tail -3 $OUT/r2_27_pytest.log
for st in 1 2 3 4; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sr_attention_fwd --csv --log-file $OUT/r2_27_launch_s$st.csv python tools/run_attn_bwd_once.py $st > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/r2_27_launch_s$st.csv")) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
for r in rows[-1:]: print("stage $st", r[ik][:40], r[iv])
PY
done
