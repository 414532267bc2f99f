#!/bin/bash
# one election per plane in the conv_tc / conv_ks MMA issuers: kernel parity + A/B against the previous library on one box
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x > $OUT/r03r_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/r03r_pytest.log | cut -c1-160
L="d0.conv1 d0.conv2 d1.conv2 d2.conv1 d2.conv2 d3.conv2 d4.conv1 u1.conv1"
for lib in new old; do
  if [ $lib == old ]; then export HCUNET_LIB=$PWD/tools/_ab/libold.so; else unset HCUNET_LIB; fi
  echo "== $lib conv"; timeout 120 python tools/kernel_bench.py conv $L 2>&1 | cut -c1-100
  echo "== $lib dgrad"; timeout 120 python tools/kernel_bench.py dgrad d0.conv2 d1.conv2 d2.conv2 d3.conv2 2>&1 | cut -c1-100
done
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-profile > $OUT/r03r_bench_$tag.json 2> $OUT/r03r_bench_$tag.err; python -c "import json;d=json.load(open('$OUT/r03r_bench_$tag.json'));print('$tag', d['ms_per_step'], d['e2e']['ms_per_step'])" || tail -5 $OUT/r03r_bench_$tag.err; }
unset HCUNET_LIB
b new A=1
b old HCUNET_LIB=$PWD/tools/_ab/libold.so
b new2 A=1
b old2 HCUNET_LIB=$PWD/tools/_ab/libold.so
timeout 300 python tools/cfg_bench.py cfg3 --steps 3 > $OUT/r03r_cfg3_new.json 2>/dev/null; python -c "import json;d=json.load(open('$OUT/r03r_cfg3_new.json'));print('cfg3 new', d['ms_per_step'], d['train_tflops'])"
HCUNET_LIB=$PWD/tools/_ab/libold.so timeout 300 python tools/cfg_bench.py cfg3 --steps 3 > $OUT/r03r_cfg3_old.json 2>/dev/null; python -c "import json;d=json.load(open('$OUT/r03r_cfg3_old.json'));print('cfg3 old', d['ms_per_step'], d['train_tflops'])"
