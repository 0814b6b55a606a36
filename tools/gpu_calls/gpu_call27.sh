#!/bin/bash
OUT=gpurun_out
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 > $OUT/s27_bench_n2.json 2> $OUT/s27_bench_n2.err ) 2> $OUT/s27_time.txt
tail -2 $OUT/s27_bench_n2.err | cut -c1-200; cat $OUT/s27_time.txt
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/s27_bench_n2.json').read().strip().splitlines()[-1])
    for k in ['value','ms_per_step','n_gpus','e2e','gpu_launches','clocks']: print(k, d[k])
except Exception as e: print("bench parse failed", e)
P
