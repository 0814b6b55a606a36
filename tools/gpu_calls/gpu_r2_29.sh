#!/bin/bash
OUT=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/r2_29_bench_n2.json 2> $OUT/r2_29_bench_n2.err; echo rc=$?
python - <<PY
import json
try:
    d=json.loads(open("$OUT/r2_29_bench_n2.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus") if k in d}, d.get("e2e"))
except Exception as e:
    print("parse failed", e)
PY
tail -5 $OUT/r2_29_bench_n2.err | cut -c1-300
