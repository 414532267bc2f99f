#!/bin/bash
OUT=gpurun_out
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r04c_bench_$tag.json 2> $OUT/r04c_bench_$tag.err; python -c "import json;d=json.load(open('$OUT/r04c_bench_$tag.json'));print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -5 $OUT/r04c_bench_$tag.err; }
b side1 HCUNET_SIDE_STREAMS=1
b side2 HCUNET_SIDE_STREAMS=2
b side1b HCUNET_SIDE_STREAMS=1
b pdl3 HCU_PDL=3
