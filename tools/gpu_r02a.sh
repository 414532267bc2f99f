#!/bin/bash
# round 2, first GPU pass: teacher-forced parity, the whole gpu suite, fp32 bench line, sanitizer on the kernel tests
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_teacher.py -q -s > $OUT/r02_teacher.log 2>&1; echo "teacher rc=$?"; grep -E "teacher-forced|passed|failed|Error|assert" $OUT/r02_teacher.log | tail -40
python -m pytest tests -m gpu -q --deselect tests/test_gpu_teacher.py > $OUT/r02_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/r02_pytest.log
python bench.py --steps 10 --warmup 3 > $OUT/r02a_bench.json 2> $OUT/r02a_bench.err; echo "bench rc=$?"; cat $OUT/r02a_bench.json
python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r02_bench_fp32.json 2> $OUT/r02_bench_fp32.err; echo "bench fp32 rc=$?"; cat $OUT/r02_bench_fp32.json
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_kernels.py -q -x > $OUT/r02_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 $OUT/r02_memcheck.log
