#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sizes.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3
( time timeout 1200 python bench.py > gpurun_out/bench_r2x.json 2> gpurun_out/bench_r2x.err ) 2> gpurun_out/bench_r2x.time; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2x.time
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2x.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step')}, d['roofline']['frac'])
print(json.dumps(d['e2e'], indent=0))
PY
tail -3 gpurun_out/bench_r2x.err
