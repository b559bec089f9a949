#!/bin/bash
# round-2 session 7: fused cluster gram-eig-update kernel (16-wide blocks): parity, timings against the three-kernel path
mkdir -p gpurun_out/r2
( timeout 900 python -m pytest tests/test_gpu_svd_dmrg.py -x -q 2>&1 | tail -15 ) > gpurun_out/r2/s7_pytest.txt
for fused in 1 0; do
  echo "== QTB_SVD_FUSED=$fused span15 D=2048"
  QTB_SVD_FUSED=$fused QTB_SVD_DEBUG=1 SVD_REPS=3 timeout 300 python profiles/svd_driver.py 15 2048 1.6 span15 2>&1 | grep -E "svd ms|lane 0|rror" | tail -16 | cut -c1-100
  echo "== QTB_SVD_FUSED=$fused decay D=4096"
  QTB_SVD_FUSED=$fused SVD_REPS=2 timeout 300 python profiles/svd_driver.py 15 4096 1.6 decay 2>&1 | grep -E "svd ms|rror"
  echo "== QTB_SVD_FUSED=$fused span15 D=4096"
  QTB_SVD_FUSED=$fused SVD_REPS=2 timeout 300 python profiles/svd_driver.py 15 4096 1.6 span15 2>&1 | grep -E "svd ms|rror"
done > gpurun_out/r2/s7_svd.txt 2>&1
QTB_PROFILE=1 timeout 900 python profiles/dmrg_sweep_bench.py 64 2048 1e-20 6 2>&1 | grep -E "profile\] sweep [345]|^sweep" > gpurun_out/r2/s7_dmrg.txt
cat gpurun_out/r2/s7_pytest.txt; grep -E "==|svd ms|rror" gpurun_out/r2/s7_svd.txt; cat gpurun_out/r2/s7_dmrg.txt
