#!/bin/bash
L="d0.conv1 d0.conv2 d1.conv1"
run() { echo "== $* $RAW"; env "$@" timeout 120 python tools/kernel_bench.py wgradrows $L $RAW 2>&1 | grep -v "^wgrad_rows:" | awk '/prof/ {n++; if (n % 8 == 0) print; next} {print}'; }
for la in 2 4 6 8 10; do RAW="" run HCU_ROWS_LA=$la; done
RAW="--raw" run HCU_ROWS_LA=10 HCU_ROWS_DEBUG=1
