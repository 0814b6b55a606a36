#!/bin/bash
# ncu evidence pass of round 1 (run under gpurun, ONE GPU): a launch list of the bench command and
# `--set full` captures of the top kernels.  Numbers printed by runs under ncu are never bench values.
set -x
OUT=gpurun_out
NCU="ncu --clock-control none"
# every launch of one eager train step (steps 3..4 of the bench command; cold-cache, serialised: compare SHARES)
timeout 900 $NCU --metrics gpu__time_duration.sum -s 40000 -c 22000 --csv --log-file $OUT/launches_bench_r1.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-graphs > $OUT/launches_bench_r1.log 2>&1
# tcgen05 attention kernels at the 1024x1024 stage-1 shape (B2, N65536, M1024, 1 head)
timeout 300 $NCU --set full --import-source on -k regex:sr_attention_fwd_kernel -s 95 -c 1 -o $OUT/prof_attn_fwd_r1 \
    python tools/bench_attention.py > $OUT/prof_attn_fwd_r1.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:sr_attention_bwd -s 186 -c 2 -o $OUT/prof_attn_bwd_r1 \
    python tools/bench_attention.py > $OUT/prof_attn_bwd_r1.log 2>&1
# HBM-bound kernels at the 1024x1024 train-step shapes
timeout 200 $NCU --set full --import-source on -k regex:local_corr_tiled -s 80 -c 1 -o $OUT/prof_local_corr_r1 \
    python tools/microbench.py --only local_corr --iters 3 > $OUT/prof_local_corr_r1.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:refine_mix -s 10 -c 1 -o $OUT/prof_refine_r1 \
    python tools/microbench.py --only refine --iters 3 > $OUT/prof_refine_r1.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:warp_bilinear_fwd -s 10 -c 1 -o $OUT/prof_warp_r1 \
    python tools/microbench.py --only warp --iters 3 > $OUT/prof_warp_r1.log 2>&1
ls -la $OUT
