#!/bin/bash
# round-2 session 1: calibrate the block SVD's knobs on the round-1 build (no code change)
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2/smi.txt 2>&1
for inner in 4 2 1; do
  echo "== QTB_SVD_INNER=$inner" 
  QTB_SVD_INNER=$inner QTB_SVD_DEBUG=1 SVD_REPS=2 timeout 300 python profiles/svd_driver.py 15 4096 1.6 decay 2>&1 | grep -E "svd ms|lane 0" | tail -40
done > gpurun_out/r2/s1_svd_inner.txt 2>&1
for lanes in 1 8; do
  echo "== QTB_SVD_LANES=$lanes"
  QTB_SVD_LANES=$lanes SVD_REPS=2 timeout 300 python profiles/svd_driver.py 15 4096 1.6 decay 2>&1 | grep -E "svd ms"
done > gpurun_out/r2/s1_svd_lanes.txt 2>&1
echo "== random (no decay)" >> gpurun_out/r2/s1_svd_lanes.txt
QTB_SVD_DEBUG=1 SVD_REPS=2 timeout 300 python profiles/svd_driver.py 15 4096 1.6 2>&1 | grep -E "svd ms|lane 0" | tail -30 >> gpurun_out/r2/s1_svd_lanes.txt
# baseline sweep timing at D=2048 (proxy for the D=4096 metric, ~40 s)
QTB_PROFILE=1 timeout 600 python profiles/dmrg_sweep_bench.py 64 2048 1e-20 6 > gpurun_out/r2/s1_dmrg_2048.txt 2>&1
tail -5 gpurun_out/r2/s1_svd_inner.txt; cat gpurun_out/r2/s1_svd_lanes.txt | grep "svd ms\|==" ; tail -12 gpurun_out/r2/s1_dmrg_2048.txt
