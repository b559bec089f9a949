#!/bin/bash
# round-2 session 8: fused kernel v2 (balanced split, register prefetch): inner sweeps 1 / 2, with QR
mkdir -p gpurun_out/r2
( timeout 900 python -m pytest tests/test_gpu_svd_dmrg.py -x -q 2>&1 | tail -5 ) > gpurun_out/r2/s8_pytest.txt
for fi in 2 1; do
  echo "== QTB_SVD_FINNER=$fi span15 D=2048"
  QTB_SVD_FINNER=$fi QTB_SVD_DEBUG=1 SVD_REPS=3 timeout 300 python profiles/svd_driver.py 15 2048 1.6 span15 2>&1 | grep -E "svd ms|lane 0|rror" | tail -3 | cut -c1-100
  echo "== QTB_SVD_FINNER=$fi span15 D=4096"
  QTB_SVD_FINNER=$fi QTB_SVD_DEBUG=1 SVD_REPS=2 timeout 300 python profiles/svd_driver.py 15 4096 1.6 span15 2>&1 | grep -E "svd ms|lane 0|rror" | tail -3 | cut -c1-100
  echo "== QTB_SVD_FINNER=$fi dmrg 2048"
  QTB_SVD_FINNER=$fi QTB_PROFILE=1 timeout 900 python profiles/dmrg_sweep_bench.py 64 2048 1e-20 6 2>&1 | grep -E "profile\] sweep [45]|^sweep 5"
done > gpurun_out/r2/s8.txt 2>&1
cat gpurun_out/r2/s8_pytest.txt gpurun_out/r2/s8.txt
