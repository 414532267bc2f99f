#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "pooled_bn or maxpool" > $OUT/r02w_pytest0.log 2>&1; rc=$?; echo "pooled rc=$rc"; tail -3 $OUT/r02w_pytest0.log
if [ $rc -ne 0 ]; then grep -E "^E " $OUT/r02w_pytest0.log | head; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02w_pytest.log
for f in 1 0; do
HCU_BN_POOLED_STATS=$f HCUNET_PROFILE_OUT=$OUT/r02w_layers_p$f.txt timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $OUT/r02w_bench_p$f.json 2> $OUT/r02w_bench_p$f.err; echo "bench pooled=$f rc=$?"
python -c "import json;d=json.load(open('$OUT/r02w_bench_p$f.json'));print('pooled=$f', d['ms_per_step'], d['e2e']['ms_per_step'], d['loss_first_last'])"
grep "bn_bwd_stats" $OUT/r02w_layers_p$f.txt | head -6 | cut -c1-100
done
