#!/bin/bash
# round-2 session 3: QR-preconditioned block Jacobi SVD: parity tests, then timing / sweep counts with and without
mkdir -p gpurun_out/r2
( timeout 900 python -m pytest tests/test_gpu_svd_dmrg.py -x -q 2>&1 | tail -15 ) > gpurun_out/r2/s3_pytest.txt
for qr in 1 0; do
  echo "== QTB_SVD_QR=$qr decay"
  QTB_SVD_QR=$qr QTB_SVD_DEBUG=1 SVD_REPS=2 timeout 300 python profiles/svd_driver.py 15 4096 1.6 decay 2>&1 | grep -E "svd ms|lane 0|rror" | tail -24 | cut -c1-90
  echo "== QTB_SVD_QR=$qr random"
  QTB_SVD_QR=$qr QTB_SVD_DEBUG=1 SVD_REPS=2 timeout 300 python profiles/svd_driver.py 15 4096 1.6 2>&1 | grep -E "svd ms|lane 0|rror" | tail -14 | cut -c1-90
done > gpurun_out/r2/s3_svd.txt 2>&1
cat gpurun_out/r2/s3_pytest.txt; grep -E "==|svd ms" gpurun_out/r2/s3_svd.txt
