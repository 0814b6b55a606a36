#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_mit_ops_gpu.py -x -q -k "dwconv or graphed" > $OUT/s3_pytest.log 2>&1; tail -4 $OUT/s3_pytest.log
timeout 120 python tools/bench_dwconv.py > $OUT/s3_dwconv_cold.log 2>&1
timeout 120 python tools/bench_dwconv.py --hot > $OUT/s3_dwconv_hot.log 2>&1
RF_DWCONV_IMPL=direct timeout 120 python tools/bench_dwconv.py --hot --only 1280 > $OUT/s3_dwconv_hot_direct.log 2>&1
NCU="ncu --clock-control none --set full --import-source on"
timeout 200 $NCU -k regex:dwconv3x3_tile -c 8 -o $OUT/s3_ncu_tile_c1280 python tools/bench_dwconv.py --only 1280 --iters 1 > $OUT/s3_ncu1.log 2>&1
timeout 200 $NCU -k regex:dwconv3x3_tile -c 8 -o $OUT/s3_ncu_tile_c256 python tools/bench_dwconv.py --only 256 --iters 1 > $OUT/s3_ncu2.log 2>&1
timeout 200 $NCU -k regex:refine -c 6 -o $OUT/s3_ncu_refine python tools/microbench.py --only refine --iters 1 > $OUT/s3_ncu3.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > $OUT/s3_bench.json 2> $OUT/s3_bench.err
cat $OUT/s3_dwconv_hot.log | cut -c1-160
python - <<'P'
import json
d=json.loads(open('gpurun_out/s3_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
P
