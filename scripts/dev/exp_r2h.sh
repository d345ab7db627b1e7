#!/bin/bash
# round-2 run H: TMA staging of the block-local column kernel, A/B against cp.async staging
mkdir -p gpurun_out
LOG=gpurun_out/exp_r2h.log
: > $LOG
echo "== correctness (TMA staging on)" >> $LOG
timeout 300 python -m pytest tests/test_gpu_toeplitz.py tests/test_gpu_sizes.py -x -q >> $LOG 2>&1
echo "rc=$?" >> $LOG
for cfg in "HIPGP_NO_TMA_COLS=1" "HIPGP_X=1"; do
  echo "== $cfg : cfg2" >> $LOG
  env $cfg timeout 200 python scripts/dev/mv_times.py pcg >> $LOG 2>&1
  echo "rc=$?" >> $LOG
  echo "== $cfg : 3-D / cfg3" >> $LOG
  env $cfg timeout 200 python scripts/dev/mv3d_times.py >> $LOG 2>&1
  echo "rc=$?" >> $LOG
done
tail -40 $LOG
