#!/bin/bash
# wgrad_rows: where does the time go?  (debug bits: 1 no MMA, 2 no transform math, 4 no dy loads, 8 no input loads, 16 spin waits)
OUT=gpurun_out
L="d0.conv2 d1.conv2 d2.conv2"
run() { echo "== $* $RAW"; env "$@" timeout 120 python tools/kernel_bench.py wgradrows $L $RAW 2>&1 | grep -v "^wgrad_rows:"; }
RAW="" run HCU_ROWS_DEBUG=0
RAW="" run HCU_ROWS_DEBUG=16
RAW="--raw" run HCU_ROWS_DEBUG=16
RAW="--raw" run HCU_ROWS_DEBUG=13
RAW="--raw" run HCU_ROWS_DEBUG=29
RAW="--raw" run HCU_ROWS_DEBUG=5
RAW="--raw" run HCU_ROWS_DEBUG=9
RAW="--raw" run HCU_ROWS_DEBUG=12
