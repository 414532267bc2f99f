#!/bin/bash
OUT=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r02g_bench_n8.json 2> $OUT/r02g_bench_n8.err; echo "bench n8 rc=$?"
python -c "import json;d=json.load(open('$OUT/r02g_bench_n8.json'));print('n8', d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['extra']['cfg4']['ms_per_step'], d['extra']['cfg5'].get('seconds_whole_stack'))" || tail -5 $OUT/r02g_bench_n8.err
