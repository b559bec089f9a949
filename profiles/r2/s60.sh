#!/bin/bash
# round-2 session 60: inner sweeps of the fused kernel's eigensolver after the column pre-sort (D=2048 sweep)
mkdir -p gpurun_out/r2
for fi in 1 2 3; do
  echo "== L=64 D=2048 QTB_SVD_FINNER=$fi" >> gpurun_out/r2/s60.txt
  QTB_SVD_FINNER=$fi QTB_PROFILE=1 timeout 900 python profiles/dmrg_sweep_bench.py 64 2048 1e-20 6 2>&1 | grep -E "profile\] sweep 5|^sweep 5" >> gpurun_out/r2/s60.txt
done
cat gpurun_out/r2/s60.txt
