#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2_18_pytest.log 2>&1; echo pytest rc=$?
grep -v "^$" $OUT/r2_18_pytest.log | tail -12 | cut -c1-1800
