#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "persistent or tcgen05" > $OUT/s22_pytest_pp.log 2>&1; PP=$?; tail -12 $OUT/s22_pytest_pp.log | cut -c1-300
timeout 120 python tools/run_gcorr_once.py time 128 2 > $OUT/s22_modes.log 2>&1
timeout 120 python tools/run_gcorr_once.py time 64 2 >> $OUT/s22_modes.log 2>&1
timeout 120 python tools/run_gcorr_once.py time 192 2 >> $OUT/s22_modes.log 2>&1
cat $OUT/s22_modes.log
timeout 100 python tools/microbench.py --iters 10 --only upsample_ce > $OUT/s22_micro.log 2>&1; grep "ce_bwd" $OUT/s22_micro.log | cut -c1-200
if [ $PP -eq 0 ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:global_corr_persist -s 4 -c 4 -o $OUT/s22_gcorr_persist \
    python tools/run_gcorr_once.py once 128 2 > $OUT/s22_ncu.log 2>&1
tail -2 $OUT/s22_ncu.log; ls -la $OUT/s22_gcorr_persist.ncu-rep
fi
