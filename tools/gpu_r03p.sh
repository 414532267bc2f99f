#!/bin/bash
OUT=gpurun_out
timeout 120 python tools/step_timeline.py > $OUT/r03p_timeline.txt 2>/dev/null; grep "^#" $OUT/r03p_timeline.txt | head -3
HCUNET_PROFILE_OUT=$OUT/r03p_layers.txt timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $OUT/r03p_bench.json 2> $OUT/r03p_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$OUT/r03p_bench.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline'])"
