#!/bin/bash
OUT=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r02g_bench_n2.json 2> $OUT/r02g_bench_n2.err; echo "bench n2 rc=$?"
python -c "import json;d=json.load(open('$OUT/r02g_bench_n2.json'));print('n2', d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['extra']['cfg4']['ms_per_step'], d['extra']['cfg5'].get('seconds_whole_stack'))" || tail -5 $OUT/r02g_bench_n2.err
timeout 300 python -m pytest tests -m gpu -q -k "parallel or overlap or allreduce" 2>&1 | tail -2
