#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "wgrad_rows" > $OUT/r04a_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 $OUT/r04a_pytest.log | cut -c1-220
