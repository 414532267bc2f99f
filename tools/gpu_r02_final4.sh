#!/bin/bash
# closing pass of the round: full suite, smoke, bench (+ per-layer profile), launch list of a bench step under ncu
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $OUT/${TAG}_pytest_gpu.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; grep "^smoke" $OUT/${TAG}_smoke.log
HCUNET_PROFILE_OUT=$OUT/${TAG}_layers.txt timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$OUT/${TAG}_bench.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['value'], d['gpu_launches'])"
timeout 120 python tools/step_timeline.py > $OUT/${TAG}_timeline.txt 2>/dev/null; grep "^#" $OUT/${TAG}_timeline.txt | head -3
timeout 120 python tools/step_breakdown.py > $OUT/${TAG}_step_breakdown.txt 2>/dev/null; head -5 $OUT/${TAG}_step_breakdown.txt
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile --no-extra > $OUT/${TAG}_plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 400 --csv --log-file $OUT/${TAG}_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile --no-extra > $OUT/${TAG}_ncu3.log 2>&1; echo "ncu list rc=$?"
