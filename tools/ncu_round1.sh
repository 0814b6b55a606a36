#!/bin/bash
# ncu evidence pass of round 1 (run under gpurun, ONE GPU): a launch list of the bench command and
# `--set full` captures of the top kernels.  Numbers printed by runs under ncu are never bench values.
OUT=gpurun_out
NCU="ncu --clock-control none"
# every launch of one eager train step (cold-cache, serialised: compare SHARES); the first 3 eager steps
# (warm-up of the per-kernel pass) are skipped by launch index
timeout 900 $NCU --metrics gpu__time_duration.sum -s 28000 -c 11000 --csv --log-file $OUT/launches_bench_r1b.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-graphs > $OUT/launches_bench_r1b.log 2>&1
# tcgen05 attention kernels at the dominant 1024x1024 stage-3 shape family (tools/bench_attention.py walks all stages)
timeout 300 $NCU --set full --import-source on -k regex:sr_attention_bwd -s 20 -c 2 -o $OUT/prof_attn_bwd_r1b \
    python tools/bench_attention.py > $OUT/prof_attn_bwd_r1b.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:sr_attention_fwd -s 10 -c 1 -o $OUT/prof_attn_fwd_r1b \
    python tools/bench_attention.py > $OUT/prof_attn_fwd_r1b.log 2>&1
# streaming kernels at the 1024x1024 train-step shapes
timeout 200 $NCU --set full --import-source on -k regex:dwconv3x3_tile -c 8 -o $OUT/prof_dwconv_tile_r1b \
    python tools/bench_dwconv.py --only 1280 --iters 1 > $OUT/prof_dwconv_tile_r1b.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:local_corr_tiled -s 80 -c 1 -o $OUT/prof_local_corr_r1b \
    python tools/microbench.py --only local_corr --iters 3 > $OUT/prof_local_corr_r1b.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:refine_mix -s 10 -c 1 -o $OUT/prof_refine_r1b \
    python tools/microbench.py --only refine --iters 3 > $OUT/prof_refine_r1b.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:warp_bilinear_fwd -s 22 -c 1 -o $OUT/prof_warp_r1b \
    python tools/microbench.py --only warp --iters 3 > $OUT/prof_warp_r1b.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:patch_embed_ln -s 4 -c 1 -o $OUT/prof_patch_embed_r1b \
    python tools/microbench.py --only patch_embed --iters 3 > $OUT/prof_patch_embed_r1b.log 2>&1
ls -la $OUT | tail -20
