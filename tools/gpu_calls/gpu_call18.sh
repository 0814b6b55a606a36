#!/bin/bash
OUT=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/s18_smoke.log 2>&1; tail -1 $OUT/s18_smoke.log
( time timeout 600 python bench.py > $OUT/s18_bench_default.json 2> $OUT/s18_bench_default.err ) 2> $OUT/s18_time.txt
tail -2 $OUT/s18_bench_default.err; cat $OUT/s18_time.txt
python - <<'P'
import json
d=json.loads(open('gpurun_out/s18_bench_default.json').read().strip().splitlines()[-1])
for k in ['value','ms_per_step','e2e','gpu_launches','roofline','cpu_baseline','corr_volume','clocks']: print(k, d[k])
P
timeout 120 python tools/microbench.py --iters 12 --only local_corr > $OUT/s18_micro.log 2>&1
grep bwd $OUT/s18_micro.log | cut -c1-230
