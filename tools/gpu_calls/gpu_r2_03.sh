#!/bin/bash
OUT=gpurun_out
for st in 1 2 3 4; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sr_attention --csv --log-file $OUT/r2_03_launch_s$st.csv python tools/run_attn_bwd_once.py $st > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/r2_03_launch_s$st.csv")) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
for r in rows[1:]: print("stage $st", r[ik][:40], r[iv])
PY
done
ncu --set full --clock-control none --import-source on -k regex:sr_attention_bwd_ws -s 2 -c 1 -o $OUT/r2_03_attn_bwd_ws_s3 python tools/run_attn_bwd_once.py 3 > $OUT/r2_03_ncu.log 2>&1; echo ncu rc=$?
