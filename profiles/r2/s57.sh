#!/bin/bash
# round-2 session 57: panel-path sweeps replayed as a CUDA graph: tests + sweeps at D = 128, 256, 512 with / without
mkdir -p gpurun_out/r2
( timeout 900 python -m pytest tests/test_gpu_svd_dmrg.py tests/test_gpu_linalg_extra.py tests/test_gpu_configs.py tests/test_gpu_sharded.py tests/test_gpu_dmrg_config0.py -x -q 2>&1 | tail -3 ) > gpurun_out/r2/s57.txt
for gr in 1 0; do
for D in 128 256 512; do
  echo "== L=64 D=$D QTB_SVD_GRAPH=$gr" >> gpurun_out/r2/s57.txt
  QTB_SVD_GRAPH=$gr QTB_PROFILE=1 timeout 600 python profiles/dmrg_sweep_bench.py 64 $D 1e-20 7 2>&1 | grep -E "profile\] sweep 6|^sweep 6" >> gpurun_out/r2/s57.txt
done
done
cat gpurun_out/r2/s57.txt
