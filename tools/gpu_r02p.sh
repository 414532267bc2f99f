#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_overlap.py tests/test_gpu_tiler.py -q -x > $OUT/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02p_pytest.log
for fl in 1 0; do
HCUNET_FLAT_ADAM=$fl timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $OUT/r02p_bench_flat$fl.json 2> $OUT/r02p_bench_flat$fl.err; echo "bench flat=$fl rc=$?"
python -c "import json;d=json.load(open('$OUT/r02p_bench_flat$fl.json'));print('flat=$fl', d['ms_per_step'], d['e2e']['ms_per_step'], d['loss_first_last'])"
done
