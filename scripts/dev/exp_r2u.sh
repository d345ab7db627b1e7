#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u.err ) 2> gpurun_out/bench_r2u.time; echo "bench rc=$?"
cat gpurun_out/bench_r2u.time | tail -4
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2u.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'].get('value'))
print(json.dumps(d.get('cpu_baselines_other_configs'), indent=0)[:3000])
PY
tail -5 gpurun_out/bench_r2u.err
