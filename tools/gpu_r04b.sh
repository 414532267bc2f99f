#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r04b_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $OUT/r04b_pytest.log | tail -2
HCUNET_PROFILE_OUT=$OUT/r04b_layers.txt timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra > $OUT/r04b_bench.json 2> $OUT/r04b_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$OUT/r04b_bench.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'])"
grep -E "up_conv" $OUT/r04b_layers.txt | grep wgrad | head -6
grep -E "cfg1: teacher" $OUT/r04b_pytest.log
