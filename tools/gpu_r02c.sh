#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_loader.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -s > $OUT/r02c_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $OUT/r02c_pytest.log | tail -3
grep -E "^[a-z0-9_]+/(fp32|mixed):|^(rich|full)" $OUT/r02c_pytest.log | head -80
grep -E "^E  |^FAILED" $OUT/r02c_pytest.log | head -40
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02c_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/r02c_smoke.log
HCUNET_PROFILE_OUT=$OUT/r02c_layers.txt python bench.py --steps 20 --warmup 5 > $OUT/r02c_bench.json 2> $OUT/r02c_bench.err; echo "bench rc=$?"; cat $OUT/r02c_bench.json; tail -5 $OUT/r02c_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02c_bench_ref.json 2>> $OUT/r02c_bench.err; cat $OUT/r02c_bench_ref.json
