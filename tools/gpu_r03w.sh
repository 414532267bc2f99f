#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_overlap.py -q -x > $OUT/r03w_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r03w_pytest.log
HCUNET_PROFILE_OUT=$OUT/r03w_layers.txt timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra > $OUT/r03w_bench.json 2> $OUT/r03w_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$OUT/r03w_bench.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'])"
grep -E "bn_bwd_apply|bn_relu_maxpool" $OUT/r03w_layers.txt | head -8
