#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests/test_gpu_svi.py -m gpu -x -q -k "slab" > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2p_tests.log
timeout 300 $TR scripts/check_slab.py 64 48 40 > gpurun_out/r2p_check_peer.log 2>&1; echo "check rc=$?" >> gpurun_out/r2p_check_peer.log
HIPGP_SLAB_CHUNKS=2 timeout 300 $TR scripts/check_slab.py 48 40 36 >> gpurun_out/r2p_check_peer.log 2>&1; echo "check rc=$?" >> gpurun_out/r2p_check_peer.log
timeout 300 $TR scripts/check_slab.py 512 512 512 bench > gpurun_out/r2p_slab_n${N}_peer.log 2>&1
for f in gpurun_out/r2p_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-600; done
