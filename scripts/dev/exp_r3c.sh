#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_toeplitz.py tests/test_gpu_bench_parity.py tests/test_gpu_api.py tests/test_gpu_sizes.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python scripts/dev/mv_times.py pcg f64 2>&1 | tail -6
