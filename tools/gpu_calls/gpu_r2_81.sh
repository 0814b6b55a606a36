#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py -x -q 2>&1 | tail -2
timeout 200 python tools/check_local_corr_tc.py 2>&1 | head -2
timeout 300 python tools/bench_gemm.py 2>/dev/null | grep total
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2_81_bench.json 2> $OUT/r2_81_bench.err; echo rc=$?
tail -1 $OUT/r2_81_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us']); print(d['corr_volume_GBps'])"
