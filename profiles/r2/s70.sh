#!/bin/bash
# round-2 session 70: compute-sanitizer memcheck over the contraction tests (16-byte staging on odd strides, edge tiles,
# device planner) and the SVD tests (pre-sort, device select, QR, fused / panel paths)
mkdir -p gpurun_out/r2
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_tensordot.py tests/test_gpu_planner.py -x -q > gpurun_out/r2/s70_tdot.txt 2>&1
echo "rc=$?" >> gpurun_out/r2/s70_tdot.txt
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_gpu_svd_dmrg.py tests/test_gpu_linalg_extra.py -x -q > gpurun_out/r2/s70_svd.txt 2>&1
echo "rc=$?" >> gpurun_out/r2/s70_svd.txt
for f in tdot svd; do echo "== $f"; grep -E "ERROR SUMMARY|Invalid|rc=|passed|failed|at .*kernel" gpurun_out/r2/s70_$f.txt | head -12; done
