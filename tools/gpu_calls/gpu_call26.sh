#!/bin/bash
# call 26: final state of the round -- full GPU suite, smoke, default bench, step profile, ncu captures of the new kernels
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/s26_pytest.log 2>&1; tail -6 $OUT/s26_pytest.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/s26_smoke.log 2>&1; tail -1 $OUT/s26_smoke.log | cut -c1-200
( time timeout 600 python bench.py > $OUT/s26_bench.json 2> $OUT/s26_bench.err ) 2> $OUT/s26_time.txt
tail -2 $OUT/s26_bench.err | cut -c1-200; cat $OUT/s26_time.txt
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/s26_bench.json').read().strip().splitlines()[-1])
    for k in ['value','ms_per_step','e2e','gpu_launches','corr_volume','clocks']: print(k, d[k])
except Exception as e: print("bench parse failed", e)
P
timeout 300 python tools/profile_step.py --out $OUT/s26_step_profile.json > $OUT/s26_profile.log 2>&1
head -30 $OUT/s26_profile.log | cut -c1-150
timeout 200 ncu --set full --clock-control none --import-source on -k regex:global_corr_persist -s 4 -c 4 -o $OUT/s26_gcorr_persist_v4 \
    python tools/run_gcorr_once.py once 128 2 > $OUT/s26_ncu_gcorr.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:upsample_ -s 3 -c 3 -o $OUT/s26_upsample_ce \
    python tools/run_ce_once.py > $OUT/s26_ncu_ce.log 2>&1
ls -la $OUT/s26_*.ncu-rep
