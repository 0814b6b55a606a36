#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests/test_bench_mode_parity_gpu.py -x -q -s > $OUT/r2_16_parity.log 2>&1; echo pytest rc=$?
grep -v "^$" $OUT/r2_16_parity.log | tail -25 | cut -c1-900
