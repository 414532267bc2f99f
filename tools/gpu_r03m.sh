#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "wgrad_rows" > $OUT/r03m_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 $OUT/r03m_pytest.log | cut -c1-180
L="d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv1 d2.conv2 d3.conv1 u1.conv1 u3.conv1"
echo "== rows"; HCU_TC_DEBUG=8 timeout 120 python tools/kernel_bench.py wgradrows $L 2>&1 | awk '/^wgrad_rows:/ {n++; if (n % 8 == 1) print; next} {print}'
echo "== old"; timeout 120 python tools/kernel_bench.py wgrad5 d2.conv2 d3.conv1 u1.conv1 2>&1
echo "== prof"; HCU_ROWS_PROF=1 timeout 120 python tools/kernel_bench.py wgradrows $L 2>&1 | awk '/prof/ {n++; if (n % 8 == 0) print; next} {print}'
