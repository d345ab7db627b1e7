#!/bin/bash
# slab v2 (bins layout) at N=2: emulated-rank tests, NCCL check, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_svi.py -m gpu -x -q -k "slab" > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2l_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/check_slab.py > gpurun_out/r2l_check.log 2>&1; echo "check rc=$?" >> gpurun_out/r2l_check.log
timeout 300 $TR scripts/check_slab.py 48 40 36 >> gpurun_out/r2l_check.log 2>&1; echo "check rc=$?" >> gpurun_out/r2l_check.log
timeout 300 $TR scripts/check_slab.py 512 512 512 bench > gpurun_out/r2l_slab_bins.log 2>&1
HIPGP_SLAB_LAYOUT=axis1 timeout 300 $TR scripts/check_slab.py 512 512 512 bench > gpurun_out/r2l_slab_axis1.log 2>&1
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2l_bench_n2.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2l_bench_n2.log
tail -3 gpurun_out/r2l_tests.log gpurun_out/r2l_check.log gpurun_out/r2l_slab_bins.log gpurun_out/r2l_slab_axis1.log
