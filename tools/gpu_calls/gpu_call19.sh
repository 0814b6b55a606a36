#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_alignment_gpu.py -x -q > $OUT/s19_pytest.log 2>&1; tail -3 $OUT/s19_pytest.log
timeout 120 python tools/microbench.py --iters 12 --only local_corr > $OUT/s19_micro.log 2>&1
grep bwd $OUT/s19_micro.log | cut -c1-200
timeout 300 python tools/time_dacs.py > $OUT/s19_dacs.log 2>&1; tail -4 $OUT/s19_dacs.log
