python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "wgrad_ws" 2>&1 | tail -5
python tools/kernel_bench.py wgradws d0.conv1 d0.conv2 d1.conv1 d1.conv2 u3.conv1 2>&1 | grep "^wgrad"
