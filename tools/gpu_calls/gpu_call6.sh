#!/bin/bash
OUT=gpurun_out
timeout 500 python -m pytest tests/test_ops_gpu.py -x -q > $OUT/s6_pytest.log 2>&1; tail -3 $OUT/s6_pytest.log
timeout 120 python tools/microbench.py --iters 20 --only local_corr > $OUT/s6_micro_lc.log 2>&1
timeout 120 python tools/microbench.py --iters 20 --only warp > $OUT/s6_micro_warp.log 2>&1
RF_WARP_IMPL=direct timeout 120 python tools/microbench.py --iters 20 --only warp > $OUT/s6_micro_warp_direct.log 2>&1
timeout 300 python tools/profile_step.py --ops --out $OUT/s6_step_profile.json > $OUT/s6_profile.log 2>&1
cut -c1-150 $OUT/s6_micro_lc.log | grep -v "64, 64, 9\|32, 32, 9"
cut -c1-170 $OUT/s6_micro_warp.log; echo; cut -c1-170 $OUT/s6_micro_warp_direct.log
