#!/bin/bash
# round-2 session 50: device block-pair matching tests + whole GPU suite
mkdir -p gpurun_out/r2
( timeout 600 python -m pytest tests/test_gpu_planner.py -x -q 2>&1 | tail -15 ) > gpurun_out/r2/s50.txt
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) >> gpurun_out/r2/s50.txt
cat gpurun_out/r2/s50.txt
