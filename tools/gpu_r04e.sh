#!/bin/bash
OUT=gpurun_out
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r04e_bench_$tag.json 2> $OUT/r04e_bench_$tag.err; python -c "import json;d=json.load(open('$OUT/r04e_bench_$tag.json'));print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -5 $OUT/r04e_bench_$tag.err; }
b base A=1
b t256s110 HCU_ROWS_IL_TMEM=256 HCU_ROWS_IL_SMEM_KB=110
b t512s110 HCU_ROWS_IL_SMEM_KB=110
b base2 A=1
b t256s200 HCU_ROWS_IL_TMEM=256
