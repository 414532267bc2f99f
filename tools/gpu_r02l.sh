#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tiler.py -q -s > $OUT/r02l_tiler.log 2>&1; echo "tiler rc=$?"; grep -E "^tiler/|passed|failed|Error|assert" $OUT/r02l_tiler.log | head -30
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02l_pytest.log
