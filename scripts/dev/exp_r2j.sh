#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/exp_r2j.log
: > $LOG
timeout 300 python -m pytest tests/test_gpu_toeplitz.py tests/test_gpu_sizes.py -x -q >> $LOG 2>&1
echo "rc=$?" >> $LOG
timeout 200 python scripts/dev/mv_times.py f64 pcg >> $LOG 2>&1
timeout 200 python scripts/dev/mv3d_times.py >> $LOG 2>&1
tail -22 $LOG
timeout 300 python scripts/prof_matvec.py > gpurun_out/plain_r2j.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cols_blk -s 2 -c 1 -f -o gpurun_out/blk_r2j python scripts/prof_matvec.py > gpurun_out/ncu_r2j.log 2>&1
echo "ncu rc=$?"
