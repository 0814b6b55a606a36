#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_multigpu_grad_equiv_gpu.py -x -q -s > $OUT/r2_30_gradeq.log 2>&1; echo rc=$?
grep -v "^$" $OUT/r2_30_gradeq.log | tail -12 | cut -c1-600
