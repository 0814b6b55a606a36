#!/bin/bash
OUT=gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/s31_smoke.log 2>&1; tail -2 $OUT/s31_smoke.log | cut -c1-250
