#!/bin/bash
# round 2, second GPU pass: full gpu suite with the new gates, smoke, A/B of the mbarrier wait hint
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q -s > $OUT/r02b_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed" $OUT/r02b_pytest.log | tail -3
grep -E "teacher-forced|^[a-z0-9_]+/(fp32|mixed):|^(rich|full)" $OUT/r02b_pytest.log | head -80
grep -E "^E  |^FAILED" $OUT/r02b_pytest.log | head -40
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02b_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/r02b_smoke.log
for v in "" nohint hint1us; do
  if [ -n "$v" ]; then export HCUNET_LIB=$PWD/hcunet_b200/libhcunet_b200.$v.so; fi
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-profile > $OUT/r02b_bench_$v.json 2> $OUT/r02b_bench_$v.err; echo "bench[$v] rc=$?"
  python -c "import json;d=json.load(open('$OUT/r02b_bench_$v.json'));print('$v', d['ms_per_step'], d['e2e']['ms_per_step'])"
  python tools/kernel_bench.py conv d0.conv1 d1.conv2 d2.conv2 d3.conv2 u1.conv1 u3.conv1 2>&1 | tail -6
done
