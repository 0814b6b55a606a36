#!/bin/bash
OUT=gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > $OUT/s9_pytest.log 2>&1; tail -3 $OUT/s9_pytest.log
timeout 300 python tools/profile_step.py --find "elementwise_kernel<128, 4,CUDAFunctor_add<c10::BFloat16>" --out $OUT/s9_step_profile.json > $OUT/s9_profile.log 2>&1
grep FIND $OUT/s9_profile.log | cut -c1-250
timeout 300 python bench.py --no-cpu-baseline > $OUT/s9_bench.json 2> $OUT/s9_bench.err
tail -3 $OUT/s9_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/s9_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['loss_src'])
P
