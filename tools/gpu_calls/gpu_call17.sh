#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/s17_pytest.log 2>&1; tail -6 $OUT/s17_pytest.log
timeout 120 python tools/microbench.py --iters 12 --only local_corr_bwd > $OUT/s17_micro.log 2>&1
grep bwd $OUT/s17_micro.log | cut -c1-230
