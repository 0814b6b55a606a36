#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/s2_pytest.log 2>&1; tail -3 $OUT/s2_pytest.log
timeout 120 python tools/microbench.py --iters 20 --only refine > $OUT/s2_micro.log 2>&1
timeout 120 python tools/microbench.py --iters 20 --only warp >> $OUT/s2_micro.log 2>&1
timeout 120 python tools/bench_dwconv.py > $OUT/s2_dwconv_cold.log 2>&1
timeout 120 python tools/bench_dwconv.py --hot > $OUT/s2_dwconv_hot.log 2>&1
NCU="ncu --clock-control none --set full --import-source on"
timeout 200 $NCU -k regex:dwconv3x3_d1_kernel -c 3 -o $OUT/s2_ncu_dw_c1280 python tools/bench_dwconv.py --only 1280 --iters 1 > $OUT/s2_ncu1.log 2>&1
timeout 200 $NCU -k regex:dwconv3x3_d1_kernel -c 3 -o $OUT/s2_ncu_dw_c256 python tools/bench_dwconv.py --only 256 --iters 1 > $OUT/s2_ncu2.log 2>&1
timeout 200 $NCU -k regex:wgrad -c 2 -o $OUT/s2_ncu_wg_c1280 python tools/bench_dwconv.py --only 1280 --iters 1 > $OUT/s2_ncu3.log 2>&1
cat $OUT/s2_micro.log | cut -c1-200
