#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "randomised" 2>&1 | tail -15 | cut -c1-250
