#!/bin/bash
# last check of the final build: full GPU suite, smoke, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r3d.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_r3d.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r3d.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/bench_r3d.json 2> gpurun_out/bench_r3d.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3d.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','matvec_ms_B16_f32_K','pcg_solve_s_B16_f32','spectrum_setup_ms_f32')}, d['e2e']['value'], d['e2e']['frac_of_value'], d['roofline']['frac'], d['gpu_launches'])
PY
tail -n 4 gpurun_out/pytest_gpu_r3d.log; tail -2 gpurun_out/smoke_r3d.log
