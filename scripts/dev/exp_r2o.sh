#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/check_slab.py 512 512 512 bench > gpurun_out/r2o_slab_n${N}_peer.log 2>&1
HIPGP_SLAB_EXCHANGE=nccl HIPGP_SLAB_CHUNKS=1 timeout 300 $TR scripts/check_slab.py 512 512 512 bench > gpurun_out/r2o_slab_n${N}_nccl.log 2>&1
for f in gpurun_out/r2o_*.log; do echo "== $f"; tail -n 2 $f | cut -c1-600; done
