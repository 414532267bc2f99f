#!/bin/bash
# ncu evidence for profiles/: launch list of a steady-state bench step + --set full captures of the top kernels.
# (run AFTER the same commands have exited 0 without ncu; numbers printed under ncu are never bench values)
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 900 -c 420 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" --launch-skip 202 -c 4 -f -o $OUT/full_conv_$TAG \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > $OUT/ncu_full1.log 2>&1; echo "ncu full conv rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"wgrad_ws_kernel|wgrad_tc5_kernel|wgrad_mma_kernel|bn_bwd_stats_h8" --launch-skip 270 -c 12 -f -o $OUT/full_bwd_$TAG \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > $OUT/ncu_full2.log 2>&1; echo "ncu full bwd rc=$?"
