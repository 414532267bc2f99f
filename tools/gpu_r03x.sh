#!/bin/bash
OUT=gpurun_out
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r03x_bench_$tag.json 2> $OUT/r03x_bench_$tag.err; python -c "import json;d=json.load(open('$OUT/r03x_bench_$tag.json'));print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -5 $OUT/r03x_bench_$tag.err; }
b il128 HCU_ROWS_IL_KB=128
b il64 HCU_ROWS_IL_KB=64
b il32 HCU_ROWS_IL_KB=32
b il128b HCU_ROWS_IL_KB=128
b il48 HCU_ROWS_IL_KB=48
