#!/bin/bash
# round-2 session 54: column pre-sort before the QR preconditioning: tests + sweeps at D = 256, 512, 2048, 4096
mkdir -p gpurun_out/r2
( timeout 900 python -m pytest tests/test_gpu_svd_dmrg.py tests/test_gpu_linalg_extra.py tests/test_gpu_configs.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -3 ) > gpurun_out/r2/s54.txt
for ps in 1 0; do
for D in 256 512; do
  echo "== L=64 D=$D QTB_SVD_PRESORT=$ps" >> gpurun_out/r2/s54.txt
  QTB_SVD_PRESORT=$ps QTB_PROFILE=1 timeout 600 python profiles/dmrg_sweep_bench.py 64 $D 1e-20 7 2>&1 | grep -E "profile\] sweep 6|^sweep 6" >> gpurun_out/r2/s54.txt
done
echo "== L=64 D=2048 QTB_SVD_PRESORT=$ps" >> gpurun_out/r2/s54.txt
QTB_SVD_PRESORT=$ps QTB_PROFILE=1 timeout 900 python profiles/dmrg_sweep_bench.py 64 2048 1e-20 6 2>&1 | grep -E "profile\] sweep 5|^sweep 5" >> gpurun_out/r2/s54.txt
done
echo "== L=100 D=4096 QTB_SVD_PRESORT=1" >> gpurun_out/r2/s54.txt
QTB_PROFILE=1 timeout 900 python profiles/dmrg_sweep_bench.py 100 4096 1e-20 7 2>&1 | grep -E "profile\] sweep [56]|^sweep [56]" >> gpurun_out/r2/s54.txt
cat gpurun_out/r2/s54.txt
