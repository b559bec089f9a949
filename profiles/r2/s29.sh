#!/bin/bash
# round-2 session 29: wall-clock split of the SVDs inside a real D=4096 DMRG sweep
mkdir -p gpurun_out/r2
QTB_SVD_DEBUG=2 timeout 600 python profiles/dmrg_sweep_bench.py 100 4096 1e-20 6 2>&1 | grep -E "phases ms|^sweep" | tail -110 > gpurun_out/r2/s29.txt
tail -45 gpurun_out/r2/s29.txt
