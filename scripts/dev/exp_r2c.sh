#!/bin/bash
# round-2 experiment C: block-local column kernel (+ shared-memory twiddle tables) vs the round-1 column kernel
mkdir -p gpurun_out
LOG=gpurun_out/exp_r2c.log
: > $LOG
echo "== correctness (block-local on)" >> $LOG
timeout 900 python -m pytest tests/test_gpu_toeplitz.py tests/test_gpu_sizes.py tests/test_gpu_api.py -x -q >> $LOG 2>&1
echo "rc=$?" >> $LOG
for cfg in "HIPGP_COLS_BLK=0" "HIPGP_COLS_BLK=1"; do
  echo "== $cfg : cfg2" >> $LOG
  env $cfg timeout 300 python scripts/dev/mv_times.py f64 pcg >> $LOG 2>&1
  echo "rc=$?" >> $LOG
  echo "== $cfg : 3-D / cfg3" >> $LOG
  env $cfg timeout 300 python scripts/dev/mv3d_times.py >> $LOG 2>&1
  echo "rc=$?" >> $LOG
done
export HIPGP_COLS_BLK=1
timeout 300 python scripts/prof_matvec.py > gpurun_out/plain_r2c.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cols_blk -s 2 -c 1 -f -o gpurun_out/blk_r2c python scripts/prof_matvec.py > gpurun_out/ncu_r2c.log 2>&1
echo "ncu rc=$?" >> $LOG
tail -50 $LOG
