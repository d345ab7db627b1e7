#!/bin/bash
# round-2 experiment B: ncu --set full of the ping-pong column kernel (token on / off)
mkdir -p gpurun_out
export HIPGP_COLS_PP=1
for ord in 1 0; do
  export HIPGP_PP_ORDER=$ord
  timeout 300 python scripts/prof_matvec.py > gpurun_out/plain_r2b_$ord.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:cols_pp -s 2 -c 1 -f -o gpurun_out/pp_r2b_ord$ord python scripts/prof_matvec.py > gpurun_out/ncu_r2b_$ord.log 2>&1
  echo "ord=$ord rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -3
