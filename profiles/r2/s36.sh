#!/bin/bash
# round-2 session 36: ncu --set full with source of the T1 grouped GEMM
mkdir -p gpurun_out/r2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 4 -c 1 -o gpurun_out/r2/s36_t1 python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/r2/s36_ncu.log 2>&1
tail -3 gpurun_out/r2/s36_ncu.log; ls -la gpurun_out/r2/s36*
