#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_overlap.py -q -x -s > $OUT/r02s_pytest0.log 2>&1; rc=$?; echo "overlap rc=$rc"; grep -E "fused BN|passed|failed|Error" $OUT/r02s_pytest0.log | head
if [ $rc -ne 0 ]; then tail -30 $OUT/r02s_pytest0.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02s_pytest.log
for f in 1 0; do
HCUNET_FUSE_BNBWD=$f HCUNET_PROFILE_OUT=$OUT/r02s_layers_f$f.txt timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $OUT/r02s_bench_f$f.json 2> $OUT/r02s_bench_f$f.err; echo "bench fuse=$f rc=$?"
python -c "import json;d=json.load(open('$OUT/r02s_bench_f$f.json'));print('fuse=$f', d['ms_per_step'], d['e2e']['ms_per_step'], d['loss_first_last'])"
done
