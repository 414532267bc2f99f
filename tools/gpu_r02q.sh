#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_overlap.py -q -x > $OUT/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02q_pytest.log
timeout 200 python tools/step_timeline.py > $OUT/r02q_timeline.txt 2> $OUT/r02q_timeline.err; echo "timeline rc=$?"; grep "^#" $OUT/r02q_timeline.txt; tail -3 $OUT/r02q_timeline.err
