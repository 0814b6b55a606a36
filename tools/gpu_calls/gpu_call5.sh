#!/bin/bash
OUT=gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/s5_pytest.log 2>&1; tail -14 $OUT/s5_pytest.log
timeout 120 python tools/bench_dwconv.py --aspp > $OUT/s5_aspp.log 2>&1
RF_DWCONV_IMPL=direct timeout 120 python tools/bench_dwconv.py --aspp > $OUT/s5_aspp_direct.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > $OUT/s5_bench.json 2> $OUT/s5_bench.err
tail -3 $OUT/s5_bench.err
cat $OUT/s5_aspp.log $OUT/s5_aspp_direct.log | cut -c1-160
python - <<'P'
import json
d=json.loads(open('gpurun_out/s5_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['loss_src'])
P
