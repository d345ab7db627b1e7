#!/bin/bash
# paired GLOBAL twiddle tables (row kernels, cols_fast): correctness + timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_toeplitz.py tests/test_gpu_sizes.py tests/test_gpu_bench_parity.py tests/test_gpu_api.py tests/test_gpu_quadform.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/dev/mv_times.py 2>&1 | tail -8
timeout 300 python scripts/dev/mv3d_times.py 2>&1 | tail -10
