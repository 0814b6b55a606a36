#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "local_corr" > $OUT/s29_pytest_lc.log 2>&1; tail -8 $OUT/s29_pytest_lc.log | cut -c1-300
timeout 200 python - > $OUT/s29_sweep.json 2> $OUT/s29_sweep.err <<'P'
import json, torch, bench
hbm = bench.peaks()[0]
print(json.dumps(bench.corr_volume_sweep(torch.device("cuda:0"), hbm)))
P
tail -2 $OUT/s29_sweep.err | cut -c1-200
python - <<'P'
import json
try:
    for c in json.loads(open('gpurun_out/s29_sweep.json').read().strip().splitlines()[-1]): print(c)
except Exception as e: print("parse failed", e)
P
