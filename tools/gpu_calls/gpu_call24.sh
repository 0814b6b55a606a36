#!/bin/bash
# call 24: warp-uniform tcgen05 / TMA issue (elect.sync) in the attention kernels and the persistent global corr
OUT=gpurun_out
timeout 300 python -m pytest tests/test_mit_ops_gpu.py -x -q -k "sr_attention" > $OUT/s24_pytest_attn.log 2>&1; AT=$?; tail -6 $OUT/s24_pytest_attn.log | cut -c1-300
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "persistent" > $OUT/s24_pytest_pp.log 2>&1; PP=$?; tail -4 $OUT/s24_pytest_pp.log | cut -c1-300
echo "AT=$AT PP=$PP"
timeout 120 python tools/run_gcorr_once.py time 128 2 > $OUT/s24_modes.log 2>&1; cat $OUT/s24_modes.log
timeout 200 python tools/bench_attention.py --fast > $OUT/s24_attn_uni1.log 2>&1; cut -c1-260 $OUT/s24_attn_uni1.log
RF_UNIFORM_ISSUE=0 timeout 200 python tools/bench_attention.py --fast > $OUT/s24_attn_uni0.log 2>&1; cut -c1-260 $OUT/s24_attn_uni0.log
if [ $AT -ne 0 ]; then export RF_UNIFORM_ISSUE=0; fi
if [ $PP -ne 0 ]; then export RF_GCORR_PERSIST=0; fi
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/s24_pytest.log 2>&1; tail -4 $OUT/s24_pytest.log | cut -c1-300
( time timeout 600 python bench.py > $OUT/s24_bench.json 2> $OUT/s24_bench.err ) 2> $OUT/s24_time.txt
tail -3 $OUT/s24_bench.err; cat $OUT/s24_time.txt
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/s24_bench.json').read().strip().splitlines()[-1])
    for k in ['value','ms_per_step','e2e','gpu_launches','roofline','corr_volume','clocks']: print(k, d[k])
    ok=d['own_kernels']
    for k in ok:
        if 'attention' in k or 'upsample_ce' in k: print(k, ok[k])
except Exception as e: print("bench parse failed", e)
P
