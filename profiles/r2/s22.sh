#!/bin/bash
mkdir -p gpurun_out/r2
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2/s22.txt 2>&1
for D in 256 512; do
  echo "== D=$D" >> gpurun_out/r2/s22.txt
  timeout 600 python profiles/dmrg_sweep_bench.py 64 $D 1e-20 7 2>&1 | grep -E "^sweep [56]" >> gpurun_out/r2/s22.txt
done
cat gpurun_out/r2/s22.txt
