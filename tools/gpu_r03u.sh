#!/bin/bash
for la in 0 3 8 10; do
echo "== HCU_ROWS_LA=$la"
HCU_ROWS_LA=$la timeout 120 python tools/kernel_bench.py wgradrows d0.conv1 d0.conv2 d1.conv1 d1.conv2 2>&1 | grep -v "^wgrad_rows:"
HCU_ROWS_LA=$la timeout 120 python tools/kernel_bench.py wgradrowsbnb d0.conv1 2>&1 | grep -v "^wgrad_rows:"
done
