#!/bin/bash
# stage attribution of conv_tc_kernel: debug bits (1 no global loads, 2 no epilogue math/stores, 4 no MMAs, 16 no stores) + role timing
OUT=gpurun_out
mkdir -p $OUT
L="d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv2 d3.conv2 u1.conv1 u3.conv1"
for dbg in 0 1 2 4 16 3 6 7; do
  echo "== HCU_TC_DEBUG=$dbg"
  HCU_TC_DEBUG=$dbg python tools/kernel_bench.py conv $L 2>&1 | grep -v Warn
done > $OUT/r02d_debugbits.txt 2>&1
echo "== raw (no BN transform)" >> $OUT/r02d_debugbits.txt
python tools/kernel_bench.py conv $L --raw >> $OUT/r02d_debugbits.txt 2>&1
export HCUNET_LIB=$PWD/hcunet_b200/libhcunet_b200.prof.so
for l in $L; do echo "== $l"; python tools/kernel_bench.py conv $l --once 2>&1 | grep -v Warn; done > $OUT/r02d_prof.txt 2>&1
cat $OUT/r02d_debugbits.txt
tail -60 $OUT/r02d_prof.txt
