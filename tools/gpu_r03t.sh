#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "flat_adam" > $OUT/r03t_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/r03t_pytest.log | cut -c1-200
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r03t_bench_$tag.json 2> $OUT/r03t_bench_$tag.err; python -c "import json;d=json.load(open('$OUT/r03t_bench_$tag.json'));print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])" || tail -5 $OUT/r03t_bench_$tag.err; }
b own1 HCUNET_OWN_ADAM=1
b own0 HCUNET_OWN_ADAM=0
b own1b HCUNET_OWN_ADAM=1
b own0b HCUNET_OWN_ADAM=0
timeout 120 python tools/step_breakdown.py 2>/dev/null | head -5
