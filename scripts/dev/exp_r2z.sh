#!/bin/bash
# final build (paired twiddle tables): bench line + ncu --set full of one plain matvec
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2z.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','matvec_ms_B16_f32_K','pcg_solve_s_B16_f32','matvec_ms_B16_f64_K','pcg_solve_s_B16_f64')}, d['e2e']['value'], d['e2e']['frac_of_value'], d['roofline']['frac'])
PY
timeout 300 python scripts/prof_matvec.py > gpurun_out/plain_r2z.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rows_fwd_fast|cols_blk|cols_fast|rows_inv_fast' -s 6 -c 3 -f -o gpurun_out/full_r2z python scripts/prof_matvec.py > gpurun_out/ncu_r2z.log 2>&1
echo "ncu rc=$?"
