#!/bin/bash
# round-2 evidence run R: ncu --set full of one plain K matvec (cfg2, B=16, fp32) with the final kernels
mkdir -p gpurun_out
timeout 300 python scripts/prof_matvec.py > gpurun_out/plain_r2r.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rows_fwd_fast|cols_blk|cols_fast|rows_inv_fast' -s 6 -c 3 -f -o gpurun_out/full_r2r python scripts/prof_matvec.py > gpurun_out/ncu_r2r.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r2r.log
timeout 300 python scripts/dev/mv_times.py 2>&1 | tail -8
timeout 300 python scripts/dev/setup_time.py 2>&1 | head -3
