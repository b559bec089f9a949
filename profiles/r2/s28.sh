#!/bin/bash
# round-2 session 28: fused-kernel phase cycles inside a real D=4096 DMRG sweep
mkdir -p gpurun_out/r2
QTB_SVD_DEBUG=3 QTB_PROFILE=1 timeout 600 python profiles/dmrg_sweep_bench.py 100 4096 1e-20 6 2>&1 | grep -E "fused phases|profile\] sweep|^sweep" | tail -44 > gpurun_out/r2/s28.txt
tail -30 gpurun_out/r2/s28.txt
