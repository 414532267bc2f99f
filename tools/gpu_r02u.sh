#!/bin/bash
OUT=gpurun_out
L="d2.conv2 d3.conv1 d3.conv2 d4.conv1 d4.conv2 u0.conv1 u1.conv1"
{
for tg in 0 5 3 2; do
  echo "== HCU_WG5_TG=$tg"; HCU_WG5_TG=$tg HCU_TC_DEBUG=8 timeout 120 python tools/kernel_bench.py wgrad5 $L 2>&1 | grep -v Warn | grep "wgrad\|wgrad_tc5:" | awk '!seen[$0]++' | cut -c1-150
done
for ns in 2 4 8 16; do
  echo "== HCU_WG5_NSEG=$ns"; HCU_WG5_NSEG=$ns timeout 120 python tools/kernel_bench.py wgrad5 $L 2>&1 | grep -v Warn | grep "^wgrad" | cut -c1-100
done
} > $OUT/r02u_wg5.txt 2>&1
cat $OUT/r02u_wg5.txt
