#!/bin/bash
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $OUT/r02t_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r02t_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02t_smoke.log 2>&1; echo "smoke rc=$?"; grep "^smoke" $OUT/r02t_smoke.log
