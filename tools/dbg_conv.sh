python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k conv_tc 2>&1 | tail -2
python tools/kernel_bench.py conv 2>&1 | grep "^conv "
