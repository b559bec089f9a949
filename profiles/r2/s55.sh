#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_svd_dmrg.py -x -q -k "large_panel" 2>&1 | tail -40 > gpurun_out/r2/s55.txt
cat gpurun_out/r2/s55.txt
