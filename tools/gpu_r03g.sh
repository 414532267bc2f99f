#!/bin/bash
L="d0.conv2 d2.conv2 u3.conv1"
run() { echo "== $* $RAW"; env "$@" timeout 120 python tools/kernel_bench.py wgradrows $L $RAW 2>&1 | grep -v "^wgrad_rows:" | awk '/prof/ {n++; if (n % 8 == 0) print; next} {print}'; }
RAW="--raw" run HCU_ROWS_PROF=1 HCU_ROWS_DEBUG=0
RAW="" run HCU_ROWS_PROF=1 HCU_ROWS_DEBUG=0
RAW="--raw" run HCU_ROWS_PROF=1 HCU_ROWS_DEBUG=77
