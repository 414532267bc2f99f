python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "wgrad" 2>&1 | tail -3
python tools/kernel_bench.py wgrad d0.conv1 d0.conv2 d1.conv1 d1.conv2 d2.conv1 u3.conv1 2>&1 | grep "^wgrad"
