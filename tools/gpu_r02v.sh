#!/bin/bash
# N GPUs: overlapped (captured) all-reduce vs the serial one; bench without the extra legs
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
for mode in 1 0; do
  HCUNET_AR_OVERLAP=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r02v_bench_n${N}_ar$mode.json 2> $OUT/r02v_bench_n${N}_ar$mode.err; echo "bench N=$N ar=$mode rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('$OUT/r02v_bench_n${N}_ar$mode.json'))
    print('N=$N ar=$mode', d['ms_per_step'], d['e2e']['ms_per_step'], d['value'])
except Exception as e: print('no json', e)
PY
done
