#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
HCUNET_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/r02m_bench.json 2> $OUT/r02m_bench.err; echo "bench rc=$?"
grep "bench rank" $OUT/r02m_bench.err | tail -8
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02m_bench.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])
print(json.dumps(d.get('extra'), indent=1)[:3000])
print(d.get('cpu_baseline'))
PY
