#!/bin/bash
# round-2 session 58: constants of the planner's tile cost model (64x64 configuration) against the T1 time
mkdir -p gpurun_out/r2
for cfg in "800 150 16" "2000 500 8" "1500 400 8" "2500 600 8" "2000 300 8" "3000 500 8" "1000 500 12" "2000 800 8"; do
  set -- $cfg
  QTB_COST_T0=$1 QTB_COST_C0=$2 QTB_COST_A=$3 timeout 300 python bench.py --steps 200 --warmup 5 --no-extra --workload T1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('T0=$1 C0=$2 A=$3 T1 ms', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4))
" >> gpurun_out/r2/s58.txt
done
cat gpurun_out/r2/s58.txt
