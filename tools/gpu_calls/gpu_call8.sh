#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_mit_ops_gpu.py -x -q -k "patch_embed or graphed" > $OUT/s8_pytest.log 2>&1; tail -3 $OUT/s8_pytest.log
timeout 120 python tools/microbench.py --iters 20 --only patch_embed > $OUT/s8_micro.log 2>&1
cut -c1-200 $OUT/s8_micro.log
timeout 300 python tools/profile_step.py --ops --out $OUT/s8_step_profile.json > $OUT/s8_profile.log 2>&1
grep -n "aten::add \|aten::add_\|aten::copy_\|aten::mul\|aten::contiguous\|aten::clone\|aten::_to_copy\|aten::cat\|aten::sum" $OUT/s8_profile.log | cut -c1-230 | head -60
