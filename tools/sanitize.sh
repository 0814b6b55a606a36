#!/bin/bash
# compute-sanitizer over the hand-rolled mbarrier / TMEM / TMA kernels (run on the GPU box):
#   memcheck  : out-of-bounds / misaligned accesses (global, shared), invalid TMA / tcgen05 operands
#   synccheck : invalid barrier usage (bar.sync with divergent / exited threads, mbarrier misuse)
#   racecheck : shared-memory hazards between threads that are not ordered by a barrier
# on the smallest shapes of the attention (forward, warp-specialised backward), GEMM, global-correlation and DACS
# tests.  Output: gpurun_out/sanitize_<tool>.log (a summary line per tool is printed).
OUT=${1:-gpurun_out}
mkdir -p $OUT
SEL='test_sr_attention_fwd[spec0] or test_sr_attention_fwd[spec5] or test_sr_attention_bwd[spec0] or test_sr_attention_bwd[spec5] or test_gemm_forward_layout[shape0] or test_gemm_forward_layout[shape3] or test_gemm_dgrad_and_wgrad_layouts[shape2] or test_dacs_mix_kernel_vs_oracle'
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --launch-timeout 120 \
    python -m pytest tests/test_mit_ops_gpu.py tests/test_gemm_gpu.py tests/test_dacs_gpu.py -x -q -k "$SEL" > $OUT/sanitize_$tool.log 2>&1
  rc=$?
  echo "sanitize $tool rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $OUT/sanitize_$tool.log | tr '\n' ' ' | cut -c1-300)"
done
