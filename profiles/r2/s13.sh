#!/bin/bash
mkdir -p gpurun_out/r2
( time timeout 1500 python -m pytest tests/test_gpu_adopt.py tests/test_gpu_configs.py tests/test_gpu_linalg_extra.py -x -q -s 2>&1 | tail -40 ) > gpurun_out/r2/s13_pytest.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2/s13_smoke.txt 2>&1
cat gpurun_out/r2/s13_pytest.txt; tail -3 gpurun_out/r2/s13_smoke.txt
