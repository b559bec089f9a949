#!/bin/bash
# round-2 session 10: fused kernel at 4 CTAs / SM
mkdir -p gpurun_out/r2
( timeout 900 python -m pytest tests/test_gpu_svd_dmrg.py -x -q 2>&1 | tail -3 ) > gpurun_out/r2/s10.txt
for D in 2048 4096; do
  echo "== span15 D=$D" >> gpurun_out/r2/s10.txt
  SVD_REPS=3 timeout 300 python profiles/svd_driver.py 15 $D 1.6 span15 2>&1 | grep -E "svd ms|rror" >> gpurun_out/r2/s10.txt
done
QTB_PROFILE=1 timeout 900 python profiles/dmrg_sweep_bench.py 64 2048 1e-20 6 2>&1 | grep -E "profile\] sweep [45]|^sweep 5" >> gpurun_out/r2/s10.txt
SVD_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,launch__cluster_max_active --clock-control none -k regex:svd_fused -s 300 -c 3 python profiles/svd_driver.py 15 2048 1.6 span15 2>&1 | grep -E "gpu__time|cluster_max" >> gpurun_out/r2/s10.txt
cat gpurun_out/r2/s10.txt
