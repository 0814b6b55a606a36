#!/bin/bash
OUT=gpurun_out
timeout 120 python tools/run_gcorr_once.py time 128 2 > $OUT/s21_modes.log 2>&1
timeout 120 python tools/run_gcorr_once.py time 128 2 64 >> $OUT/s21_modes.log 2>&1
cat $OUT/s21_modes.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:global_corr_persist -s 3 -c 3 -o $OUT/s21_gcorr_persist \
    python tools/run_gcorr_once.py once 128 2 > $OUT/s21_ncu.log 2>&1
tail -3 $OUT/s21_ncu.log; ls -la $OUT/s21_gcorr_persist.ncu-rep
