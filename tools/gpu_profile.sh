#!/bin/bash
# ncu evidence for profiles/: launch list of a steady-state bench step + --set full captures of the top kernels at the
# bench's layer shapes (tools/kernel_bench.py launches exactly one kernel per layer).  Run AFTER the same commands have
# exited 0 without ncu; numbers printed under ncu are never bench values.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 900 -c 420 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -c 6 -f -o $OUT/full_conv_$TAG \
  python tools/kernel_bench.py conv d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv2 d3.conv2 --once > $OUT/ncu_full1.log 2>&1; echo "ncu full conv rc=$?"
# K-streamed kernel: the bench model's bottom level + the classic 2D U-Net's 256 / 512 / 1024-channel levels (cfg3 shapes)
ncu --set full --clock-control none --import-source on -k regex:"conv_ks_kernel" -c 6 -f -o $OUT/full_ks_$TAG \
  python tools/kernel_bench.py conv d4.conv1 d4.conv2 c3.d2.conv2 c3.d3.conv2 c3.d4.conv2 c3.d5.conv2 --once > $OUT/ncu_full1k.log 2>&1; echo "ncu full ks rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"wgrad_ws_kernel|wgrad_mma_kernel" -c 4 -f -o $OUT/full_wgrad_$TAG \
  python tools/kernel_bench.py wgrad_auto d0.conv1 d0.conv2 d1.conv1 d1.conv2 --once > $OUT/ncu_full2.log 2>&1; echo "ncu full wgrad rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc5_kernel" -c 5 -f -o $OUT/full_wgrad5_$TAG \
  python tools/kernel_bench.py wgrad5 d2.conv2 d3.conv2 d4.conv1 c3.d3.conv2 c3.d4.conv2 --once > $OUT/ncu_full3.log 2>&1; echo "ncu full wgrad5 rc=$?"
# DRAM traffic of every tensor-core conv launch of one steady-state step (bench.py reports the per-launch mean as roofline.traffic)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_tc_kernel|conv_ks_kernel" --launch-skip 102 -c 52 --csv \
  --log-file $OUT/conv_tc_dram_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > $OUT/ncu_dram.log 2>&1; echo "ncu dram rc=$?"
# launch list of one classic-2D-U-Net train step (BASELINE config 3)
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 400 --csv --log-file $OUT/launches_cfg3_$TAG.csv \
  python tools/cfg_bench.py cfg3 --steps 1 --warmup 1 > $OUT/ncu_list3.log 2>&1; echo "ncu list cfg3 rc=$?"
