#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_bn_backward" > $OUT/r02y_pytest0.log 2>&1; rc=$?; echo "bnb kernel rc=$rc"; tail -3 $OUT/r02y_pytest0.log
if [ $rc -ne 0 ]; then grep -E "^E " $OUT/r02y_pytest0.log | head; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r02y_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02y_pytest.log
HCUNET_PROFILE_OUT=$OUT/r02y_layers.txt timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $OUT/r02y_bench.json 2> $OUT/r02y_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$OUT/r02y_bench.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['loss_first_last'], d['gpu_launches'])"
grep "bnbwd" $OUT/r02y_layers.txt | cut -c1-110
