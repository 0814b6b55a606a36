#!/bin/bash
# round 2, call 1: micro-benchmarks that decide the attention / GEMM designs + baseline tests and bench
OUT=gpurun_out
mkdir -p $OUT
timeout 120 tools/_bin/ubench_sm100 > $OUT/r2_01_ubench.json 2> $OUT/r2_01_ubench.err; echo ubench rc=$?
cat $OUT/r2_01_ubench.json
timeout 400 python -m pytest tests -m gpu -x -q > $OUT/r2_01_pytest.log 2>&1; echo pytest rc=$?
tail -3 $OUT/r2_01_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/r2_01_bench.json 2> $OUT/r2_01_bench.err; echo bench rc=$?
cut -c1-600 $OUT/r2_01_bench.json
