#!/bin/bash
mkdir -p gpurun_out/r2
( QTB_SVD_QR_PANEL=1 timeout 900 python -m pytest tests/test_gpu_svd_dmrg.py -x -q 2>&1 | tail -2 ) > gpurun_out/r2/s21.txt
for qp in 0 1; do
for D in 256 512; do
  echo "== D=$D QTB_SVD_QR_PANEL=$qp" >> gpurun_out/r2/s21.txt
  QTB_SVD_QR_PANEL=$qp QTB_PROFILE=1 timeout 600 python profiles/dmrg_sweep_bench.py 64 $D 1e-20 7 2>&1 | grep -E "profile\] sweep 6|^sweep 6" >> gpurun_out/r2/s21.txt
done
done
cat gpurun_out/r2/s21.txt
