#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 6 -c 3 -o $OUT/r2_38_gemm python tools/run_gemm_once.py 8192 320 1280 > $OUT/r2_38_ncu.log 2>&1; echo ncu rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_38_launch.csv python tools/run_gemm_once.py 8192 320 1280 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/r2_38_launch.csv")) if len(r)>5]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
for r in rows[-4:]: print(r[ik][:60], r[iv])
PY
