#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/check_slab.py 64 48 40 > gpurun_out/r2t_check.log 2>&1; echo "check rc=$?" >> gpurun_out/r2t_check.log
timeout 300 $TR scripts/check_slab.py 512 512 512 bench > gpurun_out/r2t_slab_n${N}_peer.log 2>&1
for f in gpurun_out/r2t_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-600; done
