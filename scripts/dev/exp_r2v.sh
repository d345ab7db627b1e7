#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_toeplitz.py tests/test_gpu_sizes.py tests/test_gpu_bench_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/dev/mv_times.py 2>&1 | tail -8
