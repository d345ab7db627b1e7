#!/bin/bash
# round-2 run D: full GPU test suite (incl. the drop-in tests on the staged reference), new bench line at N = 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2d.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_r2d.log
timeout 1200 python bench.py > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/bench_r2d.json; tail -5 gpurun_out/bench_r2d.err
