#!/bin/bash
# round-2 session 11: whole GPU suite on the current build, then the BASELINE metric's first half (L=100, D=4096)
mkdir -p gpurun_out/r2
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2/s11_pytest.txt 2>&1
QTB_PROFILE=1 timeout 1000 python profiles/dmrg_sweep_bench.py 100 4096 1e-20 7 2>&1 | grep -E "profile\] sweep|^sweep" > gpurun_out/r2/s11_dmrg4096.txt
cat gpurun_out/r2/s11_pytest.txt gpurun_out/r2/s11_dmrg4096.txt
