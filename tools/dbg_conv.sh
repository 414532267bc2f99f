python -m pytest tests -m gpu -q 2>&1 | tail -12
python tools/kernel_bench.py conv d0.conv1 d0.conv2 d1.conv2 2>&1 | grep "^conv "
