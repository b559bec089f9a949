#!/bin/bash
# round-2 session 27: sweep counts / group sizes of the SVDs inside a real D=4096 DMRG sweep
mkdir -p gpurun_out/r2
QTB_SVD_DEBUG=2 timeout 600 python profiles/dmrg_sweep_bench.py 100 4096 1e-20 6 2>&1 | grep -E "census|lane 0|^sweep" | cut -c1-170 | awk '/census/{print prev; print} {prev=$0} /^sweep/{print}' | tail -150 > gpurun_out/r2/s27.txt
tail -60 gpurun_out/r2/s27.txt
