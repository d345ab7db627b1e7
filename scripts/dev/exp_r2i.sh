#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_matvec.py > gpurun_out/plain_r2i.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cols_blk -s 2 -c 1 -f -o gpurun_out/blk_r2i python scripts/prof_matvec.py > gpurun_out/ncu_r2i.log 2>&1
echo "ncu rc=$?"
