#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "wgrad_rows" > $OUT/r03v_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r03v_pytest.log
timeout 120 python tools/kernel_bench.py wgradrows d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv1 d2.conv2 d3.conv1 u3.conv1 2>&1 | grep -v "^wgrad_rows:"
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r03v_bench_$tag.json 2> $OUT/r03v_bench_$tag.err; python -c "import json;d=json.load(open('$OUT/r03v_bench_$tag.json'));print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -5 $OUT/r03v_bench_$tag.err; }
b la128 A=1
b la5 HCU_ROWS_LA=5
b la128b A=1
