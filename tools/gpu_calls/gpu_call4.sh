#!/bin/bash
OUT=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > $OUT/s4_pytest.log 2>&1; tail -4 $OUT/s4_pytest.log
timeout 120 python tools/bench_dwconv.py --hot > $OUT/s4_dwconv_hot.log 2>&1
timeout 120 python tools/microbench.py --iters 20 --only warp > $OUT/s4_micro.log 2>&1
timeout 300 python tools/profile_step.py --ops --out $OUT/s4_step_profile.json > $OUT/s4_profile.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > $OUT/s4_bench.json 2> $OUT/s4_bench.err
cat $OUT/s4_dwconv_hot.log | cut -c1-160
cat $OUT/s4_micro.log | cut -c1-200
python - <<'P'
import json
d=json.loads(open('gpurun_out/s4_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
P
