#!/bin/bash
# source-level ncu capture of conv_tc_kernel on the two layers the verdict names (d0.conv2, d1.conv2) + colsum test + loader test
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_kernels.py -q -k "colsum" > $OUT/r02e_colsum.log 2>&1; echo "colsum rc=$?"; tail -3 $OUT/r02e_colsum.log
python -m pytest tests/test_gpu_loader.py -q > $OUT/r02e_loader.log 2>&1; echo "loader rc=$?"; tail -3 $OUT/r02e_loader.log
python tools/kernel_bench.py conv d0.conv2 d1.conv2 --once > $OUT/r02e_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 2 -f -o $OUT/r02e_conv \
  python tools/kernel_bench.py conv d0.conv2 d1.conv2 --once > $OUT/r02e_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/r02e_ncu.log
