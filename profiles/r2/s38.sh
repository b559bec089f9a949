#!/bin/bash
# round-2 session 38: 16-byte LDGSTS producer path (shifted layout) on the 64x64 configuration
mkdir -p gpurun_out/r2
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2/s38.txt
for v in 1 0; do
for w in T1 T2; do
  QTB_VEC16=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-extra --workload $w 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('QTB_VEC16=$v $w value', round(d['value'],3), d['unit'], 'ms', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4))
" >> gpurun_out/r2/s38.txt
done
done
cat gpurun_out/r2/s38.txt
