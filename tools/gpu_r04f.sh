#!/bin/bash
OUT=gpurun_out
HCUNET_SIDE_STREAMS=2 timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r04f_pytest.log 2>&1; echo "pytest(side2) rc=$?"; grep -E "passed|failed" $OUT/r04f_pytest.log | tail -2
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r04f_bench_$tag.json 2> $OUT/r04f_bench_$tag.err; python -c "import json;d=json.load(open('$OUT/r04f_bench_$tag.json'));print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -5 $OUT/r04f_bench_$tag.err; }
b s1a HCUNET_SIDE_STREAMS=1
b s2a HCUNET_SIDE_STREAMS=2
b s1b HCUNET_SIDE_STREAMS=1
b s2b HCUNET_SIDE_STREAMS=2
b s1c HCUNET_SIDE_STREAMS=1
b s2c HCUNET_SIDE_STREAMS=2
