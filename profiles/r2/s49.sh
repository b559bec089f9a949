#!/bin/bash
# round-2 session 49: unrolled dot and rotation loops of the panel kernel
mkdir -p gpurun_out/r2
( timeout 900 python -m pytest tests/test_gpu_svd_dmrg.py tests/test_gpu_linalg_extra.py tests/test_gpu_configs.py -x -q 2>&1 | tail -3 ) > gpurun_out/r2/s49.txt
for pc in 1; do
for D in 256 512; do
  echo "== D=$D QTB_SVD_PANEL_CROSS=$pc" >> gpurun_out/r2/s49.txt
  QTB_SVD_PANEL_CROSS=$pc QTB_PROFILE=1 timeout 600 python profiles/dmrg_sweep_bench.py 64 $D 1e-20 7 2>&1 | grep -E "profile\] sweep 6|^sweep 6" >> gpurun_out/r2/s49.txt
done
done
QTB_SVD_DEBUG=1 SVD_REPS=2 timeout 300 python profiles/svd_driver.py 11 256 1.5 span15 2>&1 | grep -E "svd ms|lane 0" | tail -4 | cut -c1-110 >> gpurun_out/r2/s49.txt
cat gpurun_out/r2/s49.txt
