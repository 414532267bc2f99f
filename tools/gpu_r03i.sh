#!/bin/bash
# wgrad_rows in the training step: parity suites, then A/B bench on one box (rows on / off / <= 32 channels)
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r03i_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 $OUT/r03i_pytest.log
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r03i_bench_$tag.json 2> $OUT/r03i_bench_$tag.err; python -c "import json;d=json.load(open('$OUT/r03i_bench_$tag.json'));print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"; }
b rows1 HCUNET_WGRADROWS=1
b rows0 HCUNET_WGRADROWS=0
b rows32 HCUNET_WGRADROWS=1 HCUNET_WGRADROWS_MAXCP=32
b rows8 HCUNET_WGRADROWS=1 HCUNET_WGRADROWS_MAXCP=8
b rows1b HCUNET_WGRADROWS=1
b rows0b HCUNET_WGRADROWS=0
