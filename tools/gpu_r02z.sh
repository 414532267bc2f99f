#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
for g in 1 0 1 0; do
HCUNET_GUARD=$g timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r02z_bench_g$g.json 2> $OUT/r02z_bench_g$g.err; python -c "import json;d=json.load(open('$OUT/r02z_bench_g$g.json'));print('guard=$g', d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"
done
