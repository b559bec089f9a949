#!/bin/bash
# round-2 session 41: ncu source view of the T1 kernel skeleton (no DMMAs, no copies)
mkdir -p gpurun_out/r2
QTB_GEMM_DEBUG=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 4 -c 1 -o gpurun_out/r2/s41_t1 python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/r2/s41_ncu.log 2>&1
ls -la gpurun_out/r2/s41*
