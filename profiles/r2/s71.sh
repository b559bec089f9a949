#!/bin/bash
# round-2 session 71: compute-sanitizer racecheck (shared-memory hazards) over smoke() and the contraction tests
mkdir -p gpurun_out/r2
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 12 python __graft_entry__.py smoke > gpurun_out/r2/s71_smoke.txt 2>&1
echo "rc=$?" >> gpurun_out/r2/s71_smoke.txt
grep -E "RACECHECK SUMMARY|hazard|rc=|smoke OK|ERROR" gpurun_out/r2/s71_smoke.txt | cut -c1-220 | head -30
