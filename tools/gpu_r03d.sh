#!/bin/bash
# round 3 (session d): first run of the row-stacked tcgen05 weight gradient (wgrad_rows.cu): parity, then isolated timings
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "wgrad_rows" > $OUT/r03d_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/r03d_pytest.log
HCU_TC_DEBUG=8 timeout 120 python tools/kernel_bench.py wgradrows d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv1 d2.conv2 u3.conv1 > $OUT/r03d_rows.txt 2>&1; echo "rows rc=$?"
cat $OUT/r03d_rows.txt
timeout 120 python tools/kernel_bench.py wgrad_auto d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv1 u3.conv1 > $OUT/r03d_auto.txt 2>&1
timeout 120 python tools/kernel_bench.py wgrad5 d2.conv2 >> $OUT/r03d_auto.txt 2>&1
cat $OUT/r03d_auto.txt
