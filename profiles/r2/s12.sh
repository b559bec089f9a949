#!/bin/bash
# round-2 session 12: fused kernel v3 (phase B restructured, runtime cluster size)
mkdir -p gpurun_out/r2
( timeout 900 python -m pytest tests/test_gpu_svd_dmrg.py -x -q 2>&1 | tail -3 ) > gpurun_out/r2/s12.txt
for D in 2048 4096; do
  for fc in 0 2 4; do
  echo "== span15 D=$D QTB_SVD_FCLUSTER=$fc" >> gpurun_out/r2/s12.txt
  QTB_SVD_FCLUSTER=$fc QTB_SVD_DEBUG=3 SVD_REPS=2 timeout 300 python profiles/svd_driver.py 15 $D 1.6 span15 2>&1 | grep -E "svd ms|rror|sweep 4: gram" | tail -3 | cut -c1-330 >> gpurun_out/r2/s12.txt
  done
done
cat gpurun_out/r2/s12.txt
