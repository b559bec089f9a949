#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 900 python -m pytest tests/test_gpu_svd_dmrg.py -x -q 2>&1 | tail -2 ) > gpurun_out/r2/s19.txt
for fl in 1 2 4; do
for D in 2048 4096; do
  echo "== span15 D=$D QTB_SVD_FLANES=$fl" >> gpurun_out/r2/s19.txt
  QTB_SVD_FLANES=$fl SVD_REPS=3 timeout 300 python profiles/svd_driver.py 15 $D 1.6 span15 2>&1 | grep -E "svd ms|rror" | tail -2 >> gpurun_out/r2/s19.txt
done
echo "== dmrg 2048 QTB_SVD_FLANES=$fl" >> gpurun_out/r2/s19.txt
QTB_SVD_FLANES=$fl QTB_PROFILE=1 timeout 900 python profiles/dmrg_sweep_bench.py 64 2048 1e-20 6 2>&1 | grep -E "profile\] sweep [5]|^sweep 5" >> gpurun_out/r2/s19.txt
done
cat gpurun_out/r2/s19.txt
