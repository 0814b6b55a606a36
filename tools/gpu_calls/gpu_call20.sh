#!/bin/bash
# call 20: persistent global-corr kernel + fused up-sampling / cross-entropy: parity, microbench, default bench
OUT=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "persistent" > $OUT/s20_pytest_pp.log 2>&1; PP=$?; tail -15 $OUT/s20_pytest_pp.log | cut -c1-300
timeout 200 python -m pytest tests/test_mit_ops_gpu.py -x -q -k "upsample" > $OUT/s20_pytest_ce.log 2>&1; CE=$?; tail -15 $OUT/s20_pytest_ce.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -x -q -k "not persistent and not upsample" > $OUT/s20_pytest.log 2>&1; tail -5 $OUT/s20_pytest.log | cut -c1-300
echo "PP=$PP CE=$CE"
if [ $PP -ne 0 ]; then export RF_GCORR_PERSIST=0; fi
ONLY="upsample_ce"
timeout 200 python tools/microbench.py --iters 10 --only upsample_ce > $OUT/s20_micro.log 2>&1
if [ $PP -eq 0 ]; then timeout 200 python tools/microbench.py --iters 10 --only global_pp >> $OUT/s20_micro.log 2>&1; fi
timeout 100 python tools/microbench.py --iters 10 --only global_tc >> $OUT/s20_micro.log 2>&1
cut -c1-260 $OUT/s20_micro.log
( time timeout 600 python bench.py > $OUT/s20_bench.json 2> $OUT/s20_bench.err ) 2> $OUT/s20_time.txt
tail -3 $OUT/s20_bench.err; cat $OUT/s20_time.txt
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/s20_bench.json').read().strip().splitlines()[-1])
    for k in ['value','ms_per_step','e2e','gpu_launches','roofline','corr_volume','clocks']: print(k, d[k])
    ok=d['own_kernels']
    for k in ok:
        if 'upsample' in k or 'global' in k: print(k, ok[k])
except Exception as e: print("bench parse failed", e)
P
