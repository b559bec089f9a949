#!/bin/bash
# round-2 session 30: lanes of the fused SVD path inside a real D=4096 DMRG sweep
mkdir -p gpurun_out/r2
for fl in 2 3; do
echo "== QTB_SVD_FLANES=$fl" >> gpurun_out/r2/s30.txt
QTB_SVD_FLANES=$fl QTB_PROFILE=1 timeout 600 python profiles/dmrg_sweep_bench.py 100 4096 1e-20 6 2>&1 | grep -E "profile\] sweep 5|^sweep 5" >> gpurun_out/r2/s30.txt
done
cat gpurun_out/r2/s30.txt
