#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2_44_pytest.log 2>&1; echo pytest rc=$?
tail -2 $OUT/r2_44_pytest.log
for own in 1 0; do
RF_OWN_GEMM=$own timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-corr-sweep > $OUT/r2_44_bench_own$own.json 2> $OUT/r2_44_bench_own$own.err; echo bench own=$own rc=$?
python - <<PY
import json
d=json.loads(open("$OUT/r2_44_bench_own$own.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches") if k in d}, d.get("e2e",{}).get("ms_per_step"))
k=d["own_kernels"]
print({n:round(v["ms_per_step"],2) for n,v in k.items() if n in ("gemm_bf16","rf_colsum","sr_attention_fwd","sr_attention_bwd")})
PY
done
