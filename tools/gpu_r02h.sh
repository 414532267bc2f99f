#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "conv_tc" > $OUT/r02h_pytest0.log 2>&1; rc=$?; echo "pytest conv_tc rc=$rc"; tail -4 $OUT/r02h_pytest0.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_teacher.py tests/test_gpu_fullsize.py -q -x > $OUT/r02h_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 $OUT/r02h_pytest.log
L="d0.conv1 d0.conv2 d1.conv1 d1.conv2 u3.conv1"
{
  echo "== train+xf"; timeout 120 python tools/kernel_bench.py conv $L 2>&1 | grep -v Warn | cut -c1-90
  echo "== train raw"; timeout 120 python tools/kernel_bench.py conv $L --raw 2>&1 | grep -v Warn | cut -c1-90
  echo "== dgrad"; timeout 120 python tools/kernel_bench.py dgrad $L 2>&1 | grep -v Warn | cut -c1-90
} > $OUT/r02h_modes.txt 2>&1
cat $OUT/r02h_modes.txt
HCUNET_PROFILE_OUT=$OUT/r02h_layers.txt timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r02h_bench.json 2> $OUT/r02h_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$OUT/r02h_bench.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_time_shares'])"
