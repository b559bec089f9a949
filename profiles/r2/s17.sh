#!/bin/bash
mkdir -p gpurun_out/r2
for cfg in "64 1" "64 0" "128 1" "128 0"; do
  set -- $cfg
  echo "== QTB_TILE=$1 QTB_BULK=$2"
  QTB_TILE=$1 QTB_BULK=$2 timeout 300 python bench.py --steps 100 --warmup 5 --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('T1 value', round(d['value'],3), d['unit'], 'ms', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4))
"
done > gpurun_out/r2/s17.txt 2>&1
cat gpurun_out/r2/s17.txt
