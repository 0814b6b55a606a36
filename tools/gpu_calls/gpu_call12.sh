#!/bin/bash
OUT=gpurun_out
( time timeout 600 python bench.py > $OUT/s12_bench_default.json 2> $OUT/s12_bench_default.err ) 2> $OUT/s12_time_default.txt
tail -c 400 $OUT/s12_bench_default.json; cat $OUT/s12_time_default.txt
( time timeout 600 python bench.py --impl reference > $OUT/s12_bench_reference.json 2> $OUT/s12_bench_reference.err ) 2> $OUT/s12_time_reference.txt
cat $OUT/s12_bench_reference.json | cut -c1-600; cat $OUT/s12_time_reference.txt
bash tools/ncu_round1.sh > $OUT/s12_ncu.log 2>&1
tail -5 $OUT/s12_ncu.log
