#!/bin/bash
# round-2 evidence run: ncu --set full captures of the new / rewritten kernels + the launch list of the bench command
OUT=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:sr_attention_bwd_ws -s 2 -c 1 -o $OUT/r02_attn_bwd_ws_s3 python tools/run_attn_bwd_once.py 3 > $OUT/r02_ncu_a.log 2>&1; echo a rc=$?
ncu --set full --clock-control none --import-source on -k regex:sr_attention_fwd_pp -s 0 -c 1 -o $OUT/r02_attn_fwd_pp_s3 python tools/run_attn_bwd_once.py 3 > $OUT/r02_ncu_b.log 2>&1; echo b rc=$?
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 6 -c 3 -o $OUT/r02_gemm_8192_320_1280 python tools/run_gemm_once.py 8192 320 1280 > $OUT/r02_ncu_c.log 2>&1; echo c rc=$?
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 4 -c 2 -o $OUT/r02_conv3x3_2x256x256_1024_256 python tools/run_conv_once.py > $OUT/r02_ncu_d.log 2>&1; echo d rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $OUT/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-corr-sweep --no-e2e > $OUT/r02_launches_bench.log 2>&1; echo launches rc=$?
wc -l $OUT/r02_launches_bench.csv
