#!/bin/bash
# round-2 evidence run Y (final): full GPU test suite, smoke, the N=1 bench line (both arms), ncu launch list of the quick bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2y.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_r2y.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r2y.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/bench_r2y.json 2> gpurun_out/bench_r2y.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_r2y.json 2> gpurun_out/bench_ref_r2y.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2y.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','spectrum_setup_ms_f32','matvec_ms_B16_f32_K','matvec_ms_B16_f32_RT','pcg_solve_s_B16_f32','matvec_ms_B16_f64_K')}, d['e2e']['value'], d['e2e']['frac_of_value'], d['roofline']['frac'], d['cpu_baseline'])
r=json.loads(open('gpurun_out/bench_ref_r2y.json').read().strip().splitlines()[-1]); print({k:r.get(k) for k in ('value','ms_per_step','impl')}, r.get('cpu_baseline'))
PY
timeout 300 python bench.py --steps 2 --warmup 3 --quick --no-cpu > gpurun_out/plain2_r2y.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/launches_r2y.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu > gpurun_out/ncu2_r2y.log 2>&1
echo "ncu launches rc=$?"
tail -n 5 gpurun_out/pytest_gpu_r2y.log
