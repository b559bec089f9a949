#!/bin/bash
# round-2 session 61: CTAs per block pair of the fused SVD kernel at D=4096 (31-sector synthetic theta), final build
mkdir -p gpurun_out/r2
for fc in 0 1 2 4; do
  echo "== QTB_SVD_FCLUSTER=$fc (0 = automatic)" >> gpurun_out/r2/s61.txt
  QTB_SVD_FCLUSTER=$fc SVD_REPS=3 timeout 300 python profiles/svd_driver.py 31 4096 3.2 span15 2>&1 | grep -E "svd ms" | tail -2 >> gpurun_out/r2/s61.txt
done
cat gpurun_out/r2/s61.txt
