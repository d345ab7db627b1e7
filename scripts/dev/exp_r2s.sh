#!/bin/bash
# bench at N GPUs with NUMA-local pinned buffers; host topology probe
N=${1:-2}
mkdir -p gpurun_out
(lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name"; nvidia-smi topo -m 2>/dev/null | head -14; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node)"; fi; done) > gpurun_out/r2s_topo_n$N.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2s_bench_n$N.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2s_bench_n$N.log
cat gpurun_out/r2s_topo_n$N.log | head -30
python - <<PY
import json
L=[l for l in open("gpurun_out/r2s_bench_n$N.log").read().splitlines() if l.startswith("{")]
d=json.loads(L[-1]); print(d["value"], d["e2e"]); print(json.dumps(d.get("slab_cfg5"))[:1200]); print(json.dumps(d.get("svi_cfg3"))[:600])
PY
