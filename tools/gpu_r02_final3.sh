#!/bin/bash
# last validation of the round: full GPU suite, smoke, default bench on one GPU, then the same bench on two GPUs (torchrun)
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $OUT/${TAG}_pytest_gpu.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; grep "^smoke" $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_default.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['value'], d['gpu_launches'])"
if [ "$2" == "n2" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err; echo "bench n2 rc=$?"
python -c "import json;d=json.load(open('$OUT/${TAG}_bench_n2.json'));print('n2', d['ms_per_step'], d['e2e']['ms_per_step'], d['value'])" || tail -5 $OUT/${TAG}_bench_n2.err
fi
