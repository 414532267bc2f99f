#!/bin/bash
OUT=gpurun_out
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r04h_bench_$tag.json 2> $OUT/r04h_bench_$tag.err; python -c "import json;d=json.load(open('$OUT/r04h_bench_$tag.json'));print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -5 $OUT/r04h_bench_$tag.err; }
b ms16 HCU_ROWS_MINSTEPS=16
b ms8 HCU_ROWS_MINSTEPS=8
b ms32 HCU_ROWS_MINSTEPS=32
b ms16b HCU_ROWS_MINSTEPS=16
b ms24 HCU_ROWS_MINSTEPS=24
