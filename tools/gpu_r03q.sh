#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "wgrad_rows" > $OUT/r03q_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/r03q_pytest.log | cut -c1-180
L="d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv1 d2.conv2 d3.conv1 u1.conv1 u3.conv1"
echo "== rows"; HCU_TC_DEBUG=8 timeout 120 python tools/kernel_bench.py wgradrows $L 2>&1 | awk '/^wgrad_rows:/ {n++; if (n % 8 == 1) print; next} {print}' | cut -c1-200
timeout 120 python tools/step_timeline.py > $OUT/r03q_timeline.txt 2>/dev/null; grep "^#" $OUT/r03q_timeline.txt | head -3
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r03q_bench_$tag.json 2> $OUT/r03q_bench_$tag.err; python -c "import json;d=json.load(open('$OUT/r03q_bench_$tag.json'));print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -5 $OUT/r03q_bench_$tag.err; }
b cp32o32 HCUNET_WGRADROWS_MAXCP=32 HCUNET_WGRADROWS_MAXCP_OUT=32
b cp32o64 HCUNET_WGRADROWS_MAXCP=32 HCUNET_WGRADROWS_MAXCP_OUT=64
