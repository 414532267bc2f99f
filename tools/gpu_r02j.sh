#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
export HCUNET_LIB=$PWD/hcunet_b200/libhcunet_b200.prof.so
{
echo "== PROF d0.conv2 train+xf"; timeout 60 python tools/kernel_bench.py conv d0.conv2 --once 2>&1 | grep prof
echo "== PROF d0.conv2 train raw (bulk)"; timeout 60 python tools/kernel_bench.py conv d0.conv2 --raw --once 2>&1 | grep prof
echo "== PROF d0.conv2 dgrad (bulk)"; timeout 60 python tools/kernel_bench.py dgrad d0.conv2 --once 2>&1 | grep prof
echo "== PROF d0.conv2 dgrad (no bulk)"; HCU_TC_BULK=0 timeout 60 python tools/kernel_bench.py dgrad d0.conv2 --once 2>&1 | grep prof
echo "== PROF d1.conv2 train+xf"; timeout 60 python tools/kernel_bench.py conv d1.conv2 --once 2>&1 | grep prof
echo "== PROF u3.conv1 train raw"; timeout 60 python tools/kernel_bench.py conv u3.conv1 --raw --once 2>&1 | grep prof
} > $OUT/r02j_prof.txt 2>&1
cat $OUT/r02j_prof.txt
unset HCUNET_LIB
echo "== dgrad nobulk"; HCU_TC_BULK=0 timeout 120 python tools/kernel_bench.py dgrad d0.conv1 d0.conv2 d1.conv1 2>&1 | grep -v Warn | cut -c1-90
echo "== raw nobulk"; HCU_TC_BULK=0 timeout 120 python tools/kernel_bench.py conv d0.conv1 d0.conv2 d1.conv1 --raw 2>&1 | grep -v Warn | cut -c1-90
