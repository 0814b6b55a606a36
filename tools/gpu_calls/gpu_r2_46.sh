#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_mit_ops_gpu.py -x -q -k "upsample_bilinear" 2>&1 | tail -2
bash tools/sanitize.sh $OUT
