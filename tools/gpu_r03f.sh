#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "wgrad_rows" > $OUT/r03f_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $OUT/r03f_pytest.log
L="d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv1 d2.conv2 u3.conv1"
run() { echo "== $* $RAW"; env "$@" timeout 120 python tools/kernel_bench.py wgradrows $L $RAW 2>&1 | grep -v "^wgrad_rows:" | awk '/prof/ {n++; if (n % 8 == 0) print; next} {print}'; }
RAW="" run HCU_ROWS_DEBUG=0
RAW="--raw" run HCU_ROWS_DEBUG=1
RAW="" run HCU_ROWS_PROF=1
