#!/bin/bash
# round-2 session 72: the sharded tile rule (64x64 below four 128x128 tiles per SM) applied on one GPU too? D=2048 sweep
mkdir -p gpurun_out/r2
for r in 1 0 1; do
  echo "== L=64 D=2048 QTB_TILE_RULE=$r" >> gpurun_out/r2/s72b.txt
  QTB_TILE_RULE=$r QTB_PROFILE=1 timeout 300 python profiles/dmrg_sweep_bench.py 64 2048 1e-20 6 2>&1 | grep -E "profile\] sweep 5|^sweep 5" >> gpurun_out/r2/s72b.txt
done
cat gpurun_out/r2/s72b.txt
