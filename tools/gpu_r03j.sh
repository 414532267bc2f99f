#!/bin/bash
OUT=gpurun_out
timeout 120 python tools/step_timeline.py > $OUT/r03j_timeline_rows.txt 2>/dev/null; grep "^#" $OUT/r03j_timeline_rows.txt | head -3
HCU_ROWS_TMEM=256 timeout 120 python tools/step_timeline.py > $OUT/r03j_timeline_rows256.txt 2>/dev/null; grep "^#" $OUT/r03j_timeline_rows256.txt | head -3
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r03j_bench_$tag.json 2> $OUT/r03j_bench_$tag.err; python -c "import json;d=json.load(open('$OUT/r03j_bench_$tag.json'));print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'])"; }
b t512 HCU_ROWS_TMEM=512
b t256 HCU_ROWS_TMEM=256
