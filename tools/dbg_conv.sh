python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in 0 1; do echo "wide=$w"; HCU_TC_WIDE=$w python tools/kernel_bench.py conv d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv1 u3.conv1 2>&1 | grep "^conv "; done
HCU_TC_DEBUG=8 python tools/kernel_bench.py conv d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv1 u3.conv1 --once 2>&1 | grep conv_tc
