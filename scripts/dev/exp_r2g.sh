#!/bin/bash
# round-2 run G: drop-in tests (incl. bidiag), bench line, ncu launch list + full capture of one plain K matvec (traffic)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_api.py -q -x 2>&1 | tail -4
timeout 300 python scripts/dev/setup_time.py 2>&1 | head -2
timeout 1200 python bench.py --no-multi > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2g.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','spectrum_setup_ms_f32','matvec_ms_B16_f32_K','pcg_solve_s_B16_f32')}, d['e2e']['value'], d['e2e']['frac_of_value'], d['roofline']['frac'], d['cpu_baseline'])
PY
timeout 300 python scripts/prof_matvec.py > gpurun_out/plain_r2g.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rows_fwd_fast|cols_blk|cols_fast|rows_inv_fast' -s 6 -c 3 -f -o gpurun_out/full_r2g python scripts/prof_matvec.py > gpurun_out/ncu_r2g.log 2>&1
echo "ncu rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --quick --no-cpu > gpurun_out/plain2_r2g.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/launches_r2g.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu > gpurun_out/ncu2_r2g.log 2>&1
echo "ncu launches rc=$?"
