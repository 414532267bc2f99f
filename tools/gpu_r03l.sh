#!/bin/bash
run() { echo "== $*"; env "$@" 2>&1 | grep -v "^wgrad_rows:" | awk '/prof/ {n++; if (n % 8 == 0) print; next} {print}'; }
run timeout 120 python tools/kernel_bench.py wgradrowsbnb d0.conv1 d0.conv2
run timeout 120 python tools/kernel_bench.py wgradrows d0.conv1 d0.conv2 --raw
run env HCU_ROWS_PROF=1 timeout 120 python tools/kernel_bench.py wgradrowsbnb d0.conv1 d0.conv2
run env HCU_ROWS_DEBUG=1 timeout 120 python tools/kernel_bench.py wgradrowsbnb d0.conv1 d0.conv2
