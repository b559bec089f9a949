#!/bin/bash
mkdir -p gpurun_out/r2
( timeout 900 python -m pytest tests/test_gpu_svd_dmrg.py tests/test_gpu_linalg_extra.py -x -q 2>&1 | tail -3 ) > gpurun_out/r2/s18.txt
for D in 2048 4096; do
  echo "== span15 D=$D" >> gpurun_out/r2/s18.txt
  QTB_SVD_DEBUG=3 SVD_REPS=3 timeout 300 python profiles/svd_driver.py 15 $D 1.6 span15 2>&1 | grep -E "svd ms|rror|sweep 4: gram" | tail -4 | cut -c1-330 >> gpurun_out/r2/s18.txt
done
QTB_PROFILE=1 timeout 900 python profiles/dmrg_sweep_bench.py 64 2048 1e-20 6 2>&1 | grep -E "profile\] sweep [45]|^sweep 5" >> gpurun_out/r2/s18.txt
cat gpurun_out/r2/s18.txt
