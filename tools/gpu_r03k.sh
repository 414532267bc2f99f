#!/bin/bash
# fused BatchNorm-backward apply in the first layer's weight gradient: parity + A/B bench on one box
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "wgrad_rows or fullsize or parity" > $OUT/r03k_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/r03k_pytest.log | cut -c1-200
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r03k_bench_$tag.json 2> $OUT/r03k_bench_$tag.err; python -c "import json;d=json.load(open('$OUT/r03k_bench_$tag.json'));print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -5 $OUT/r03k_bench_$tag.err; }
b fuse1 HCUNET_FUSE_APPLY=1
b fuse0 HCUNET_FUSE_APPLY=0
b fuse1b HCUNET_FUSE_APPLY=1
b fuse0b HCUNET_FUSE_APPLY=0
timeout 120 python tools/step_timeline.py > $OUT/r03k_timeline.txt 2>/dev/null; grep "^#" $OUT/r03k_timeline.txt | head -3; tail -9 $OUT/r03k_timeline.txt | cut -c1-110
