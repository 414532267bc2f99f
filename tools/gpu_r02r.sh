#!/bin/bash
OUT=gpurun_out
L="d2.conv1 d2.conv2 d3.conv1 d3.conv2 u0.conv1 u1.conv1"
{
for h in 1 2; do
echo "== train hint=$h"; timeout 120 python tools/kernel_bench.py conv $L --hint=$h 2>&1 | grep -v Warn | cut -c1-150
echo "== dgrad hint=$h"; timeout 120 python tools/kernel_bench.py dgrad $L --hint=$h 2>&1 | grep -v Warn | cut -c1-150
done
} > $OUT/r02r_hint.txt 2>&1
cat $OUT/r02r_hint.txt
