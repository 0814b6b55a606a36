#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/s15_pytest.log 2>&1; tail -4 $OUT/s15_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $OUT/s15_bench.json 2> $OUT/s15_bench.err
tail -3 $OUT/s15_bench.err
RF_ATTN_BWD_SERIAL=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e > $OUT/s15_bench_serial.json 2> $OUT/s15_bench_serial.err
python - <<'P'
import json
for f in ['gpurun_out/s15_bench.json','gpurun_out/s15_bench_serial.json']:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e'] and d['e2e']['value'], d['gpu_launches'], d['loss_src'], d['roofline']['kernel'], d['roofline']['frac'])
P
