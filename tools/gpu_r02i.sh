#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python -m pytest tests/test_gpu_kernels.py -q -x -k "colsum or conv_tc" > $OUT/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r02i_pytest.log
{ timeout 200 python tools/step_breakdown.py; HCUNET_OVERLAP=0 timeout 200 python tools/step_breakdown.py; } 2>&1 | grep overlap | tee $OUT/r02i_breakdown.txt
HCUNET_PROFILE_OUT=$OUT/r02i_layers.txt timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r02i_bench.json 2> $OUT/r02i_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$OUT/r02i_bench.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_time_shares'])"
