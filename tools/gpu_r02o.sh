#!/bin/bash
# 2 GPUs: overlapped (captured) all-reduce vs the serial one
OUT=gpurun_out
mkdir -p $OUT
for mode in 1; do
  HCUNET_AR_OVERLAP=$mode HCUNET_BENCH_VERBOSE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r02o_bench_n2_ar$mode.json 2> $OUT/r02o_bench_n2_ar$mode.err; echo "bench ar=$mode rc=$?"
  tail -3 $OUT/r02o_bench_n2_ar$mode.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open('$OUT/r02o_bench_n2_ar$mode.json'))
    print('ar=$mode', d['ms_per_step'], d['e2e']['ms_per_step'], d['config'].get('gradient_exchange','')[:60], {k:(v.get('ms_per_step') or v.get('seconds_whole_stack') or v) for k,v in (d.get('extra') or {}).items()})
except Exception as e: print('no json', e)
PY
done
