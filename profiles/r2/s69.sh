#!/bin/bash
# round-2 session 69: compute-sanitizer memcheck over smoke() (every kernel family once, small sizes)
mkdir -p gpurun_out/r2
timeout 560 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2/s69.txt 2>&1
echo "rc=$?" >> gpurun_out/r2/s69.txt
grep -E "ERROR SUMMARY|Invalid|rc=|smoke OK|at .*kernel" gpurun_out/r2/s69.txt | head -30
