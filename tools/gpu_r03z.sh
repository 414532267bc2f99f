#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_gpu_loader.py -q 2>&1 | tail -2
HCUNET_PROFILE_OUT=$OUT/r03z_layers.txt timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra > $OUT/r03z_bench.json 2> $OUT/r03z_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$OUT/r03z_bench.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'])"
grep -E "load_stack" $OUT/r03z_layers.txt | head -3
