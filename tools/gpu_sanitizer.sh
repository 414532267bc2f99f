#!/bin/bash
# usage: bash tools/gpu_sanitizer.sh memcheck|racecheck   (ONE tool per gpurun call, B200_PROFILING.md)
TOOL=${1:-memcheck}
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tiler.py -q -x > $OUT/r02_san_plain.log 2>&1; rc=$?; echo "plain rc=$rc"; tail -2 $OUT/r02_san_plain.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tiler.py -q -x \
  > $OUT/r02_sanitizer_$TOOL.log 2>&1; echo "$TOOL rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|hazard" $OUT/r02_sanitizer_$TOOL.log | head -20
