#!/bin/bash
OUT=gpurun_out
for pr in 0 1; do
RF_GEMM_PAIRS=$pr timeout 60 tools/_bin/trace_gemm 8192 1280 320 > $OUT/r2_80_trace_pairs$pr.txt; head -1 $OUT/r2_80_trace_pairs$pr.txt; grep "block 0 warp 0:\|block 0 warp 9:" $OUT/r2_80_trace_pairs$pr.txt | cut -c1-420
RF_GEMM_PAIRS=$pr timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -1
RF_GEMM_PAIRS=$pr timeout 300 python tools/bench_gemm.py 2>/dev/null > $OUT/r2_80_bench_gemm_pairs$pr.jsonl; grep total $OUT/r2_80_bench_gemm_pairs$pr.jsonl
done
