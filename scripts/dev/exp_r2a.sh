#!/bin/bash
# round-2 experiment A: ping-pong column kernel vs the single-group kernel (A/B through environment switches)
mkdir -p gpurun_out
LOG=gpurun_out/exp_r2a.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> $LOG 2>&1
echo "== correctness (ping-pong on, token on)" >> $LOG
timeout 600 python -m pytest tests/test_gpu_toeplitz.py tests/test_gpu_sizes.py -x -q >> $LOG 2>&1
echo "rc=$?" >> $LOG
for cfg in "HIPGP_COLS_PP=0" "HIPGP_COLS_PP=1 HIPGP_PP_ORDER=0" "HIPGP_COLS_PP=1 HIPGP_PP_ORDER=1"; do
  echo "== $cfg : cfg2" >> $LOG
  env $cfg timeout 300 python scripts/dev/mv_times.py f64 pcg >> $LOG 2>&1
  echo "rc=$?" >> $LOG
  echo "== $cfg : 3-D / cfg3" >> $LOG
  env $cfg timeout 300 python scripts/dev/mv3d_times.py >> $LOG 2>&1
  echo "rc=$?" >> $LOG
done
tail -60 $LOG
