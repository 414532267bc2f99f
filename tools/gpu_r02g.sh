#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
L="d0.conv1 d0.conv2 d1.conv1 d1.conv2 u3.conv1"
{
for env in "" "HCU_TC_BULK=0" "HCU_TC_SPEC=0"; do
  echo "== train+xf $env"; env $env python tools/kernel_bench.py conv $L 2>&1 | grep -v Warn | cut -c1-90
  echo "== train raw $env"; env $env python tools/kernel_bench.py conv $L --raw 2>&1 | grep -v Warn | cut -c1-90
  echo "== dgrad $env"; env $env python tools/kernel_bench.py dgrad $L 2>&1 | grep -v Warn | cut -c1-90
done
} > $OUT/r02g_modes.txt 2>&1
cat $OUT/r02g_modes.txt
export HCUNET_LIB=$PWD/hcunet_b200/libhcunet_b200.prof.so
{
echo "== PROF d0.conv2 train+xf"; python tools/kernel_bench.py conv d0.conv2 --once 2>&1 | grep prof
echo "== PROF d0.conv1 train raw (bulk)"; python tools/kernel_bench.py conv d0.conv1 --raw --once 2>&1 | grep prof
echo "== PROF d0.conv2 dgrad (bulk)"; python tools/kernel_bench.py dgrad d0.conv2 --once 2>&1 | grep prof
echo "== PROF d0.conv2 dgrad (no bulk)"; HCU_TC_BULK=0 python tools/kernel_bench.py dgrad d0.conv2 --once 2>&1 | grep prof
} > $OUT/r02g_prof.txt 2>&1
cat $OUT/r02g_prof.txt
