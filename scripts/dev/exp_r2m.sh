#!/bin/bash
# slab bins layout with chunked, overlapped exchange: correctness + timing vs number of chunks (N = WORLD GPUs)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_gpu_svi.py -m gpu -x -q -k "slab" > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests.log
for c in 2 4; do HIPGP_SLAB_CHUNKS=$c timeout 300 $TR scripts/check_slab.py 64 48 40 > gpurun_out/r2m_check_c$c.log 2>&1; echo "check rc=$?" >> gpurun_out/r2m_check_c$c.log; done
for c in 1 2 4 8; do HIPGP_SLAB_CHUNKS=$c timeout 300 $TR scripts/check_slab.py 512 512 512 bench > gpurun_out/r2m_slab_n${N}_c$c.log 2>&1; done
HIPGP_SLAB_LAYOUT=axis1 timeout 300 $TR scripts/check_slab.py 512 512 512 bench > gpurun_out/r2m_slab_n${N}_axis1.log 2>&1
for f in gpurun_out/r2m_*.log; do echo "== $f"; tail -n 2 $f | cut -c1-300; done
