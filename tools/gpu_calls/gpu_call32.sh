#!/bin/bash
OUT=gpurun_out
timeout 42 python tools/bench_hrda.py --steps 2 --warmup 1 > $OUT/s32_hrda.json 2> $OUT/s32_hrda.err; echo rc=$?
tail -c 900 $OUT/s32_hrda.json; tail -3 $OUT/s32_hrda.err | cut -c1-300
