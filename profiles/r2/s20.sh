#!/bin/bash
mkdir -p gpurun_out/r2
for pr in 879 400 200 100; do
for D in 256 512; do
  echo "== D=$D QTB_SVD_PANEL_ROWS=$pr" >> gpurun_out/r2/s20.txt
  QTB_SVD_PANEL_ROWS=$pr QTB_PROFILE=1 timeout 600 python profiles/dmrg_sweep_bench.py 64 $D 1e-20 7 2>&1 | grep -E "profile\] sweep 6|^sweep 6" >> gpurun_out/r2/s20.txt
done
done
cat gpurun_out/r2/s20.txt
