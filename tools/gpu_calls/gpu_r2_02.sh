#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_mit_ops_gpu.py -x -q -k "sr_attention" > $OUT/r2_02_pytest.log 2>&1; echo pytest rc=$?
tail -15 $OUT/r2_02_pytest.log
RF_ATTN_BWD=old timeout 200 python tools/bench_attention.py --fast > $OUT/r2_02_attn_old.jsonl 2>$OUT/r2_02_attn_old.err; echo rc=$?
cat $OUT/r2_02_attn_old.jsonl | cut -c1-400
timeout 200 python tools/bench_attention.py --fast > $OUT/r2_02_attn_ws.jsonl 2>$OUT/r2_02_attn_ws.err; echo rc=$?
cat $OUT/r2_02_attn_ws.jsonl | cut -c1-400; tail -3 $OUT/r2_02_attn_ws.err
