#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q > $OUT/r2_35_gemm.log 2>&1; echo gemm rc=$?
grep -v "^$" $OUT/r2_35_gemm.log | tail -12 | cut -c1-300
timeout 300 python tools/bench_gemm.py > $OUT/r2_35_bench_gemm.jsonl 2> $OUT/r2_35_bench_gemm.err; echo rc=$?
cat $OUT/r2_35_bench_gemm.jsonl | cut -c1-330; tail -2 $OUT/r2_35_bench_gemm.err
