python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k wgrad 2>&1 | tail -2
python tools/kernel_bench.py wgrad d2.conv1 2>&1 | grep "^wgrad"
