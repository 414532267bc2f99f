#!/bin/bash
# specialised conv_tc variants: kernel + parity tests, per-layer timing, bench step
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_teacher.py -q -x > $OUT/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02f_pytest.log
L="d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv2 d3.conv2 u1.conv1 u3.conv1"
python tools/kernel_bench.py conv $L 2>&1 | grep -v Warn | cut -c1-120 | tee $OUT/r02f_layers_spec.txt
HCU_TC_BULK=0 python tools/kernel_bench.py conv $L 2>&1 | grep -v Warn | cut -c1-120 | tee $OUT/r02f_layers_nobulk.txt
HCU_TC_SPEC=0 python tools/kernel_bench.py conv $L 2>&1 | grep -v Warn | cut -c1-120 | tee $OUT/r02f_layers_generic.txt
HCUNET_PROFILE_OUT=$OUT/r02f_layers.txt python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/r02f_bench.json 2> $OUT/r02f_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$OUT/r02f_bench.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_time_shares'])"
