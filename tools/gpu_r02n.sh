#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_overlap.py tests/test_gpu_parity.py -q -x > $OUT/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r02n_pytest.log
