#!/bin/bash
OUT=gpurun_out
for ns in 1 2; do
HCUNET_SIDE_STREAMS=$ns timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r02x_bench_s$ns.json 2> $OUT/r02x_bench_s$ns.err; echo "bench side=$ns rc=$?"
python -c "import json;d=json.load(open('$OUT/r02x_bench_s$ns.json'));print('side=$ns', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
HCU_PDL=3 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r02x_bench_pdl3.json 2> $OUT/r02x_bench_pdl3.err; python -c "import json;d=json.load(open('$OUT/r02x_bench_pdl3.json'));print('pdl=3', d['ms_per_step'], d['e2e']['ms_per_step'])"
