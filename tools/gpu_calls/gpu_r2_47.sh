#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q -k "conv3x3" > $OUT/r2_47_conv.log 2>&1; echo conv rc=$?
grep -v "^$" $OUT/r2_47_conv.log | tail -14 | cut -c1-300
