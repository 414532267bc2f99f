#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
export HCUNET_LIB=$PWD/hcunet_b200/libhcunet_b200.prof.so
{
echo "== CTA timeline d0.conv2 dgrad (bulk)"; timeout 60 python tools/kernel_bench.py dgrad d0.conv2 --once 2>&1 | grep "cta" | sort -k3 -n
echo "== CTA timeline u3.conv1 train raw"; timeout 60 python tools/kernel_bench.py conv u3.conv1 --raw --once 2>&1 | grep "cta" | sort -k3 -n
echo "== CTA timeline d2.conv2 train xf (generic)"; timeout 60 python tools/kernel_bench.py conv d2.conv2 --once 2>&1 | grep "cta" | sort -k3 -n
} > $OUT/r02k_cta.txt 2>&1
cat $OUT/r02k_cta.txt
