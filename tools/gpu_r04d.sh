#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "flat_adam" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-profile > gpurun_out/r04d_bench.json 2> gpurun_out/r04d_bench.err; python -c "import json;d=json.load(open('gpurun_out/r04d_bench.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'])"
