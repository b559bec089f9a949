#!/bin/bash
# round-2 session 45: full GPU suite + default bench (all workloads) + T1 with 128x128 tiles
mkdir -p gpurun_out/r2
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2/s45_pytest.txt 2>&1
( time timeout 1500 python bench.py ) > gpurun_out/r2/s45_bench.json 2> gpurun_out/r2/s45_bench.err
QTB_TILE=128 timeout 300 python bench.py --steps 100 --warmup 5 --no-extra --workload T1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('QTB_TILE=128 T1 ms', round(d['ms_per_step'],5))
" > gpurun_out/r2/s45_t128.txt
cat gpurun_out/r2/s45_pytest.txt gpurun_out/r2/s45_t128.txt; tail -5 gpurun_out/r2/s45_bench.err; cat gpurun_out/r2/s45_bench.json
