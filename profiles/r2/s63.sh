#!/bin/bash
# round-2 session 63: the 64x64 configuration on the large-K workload (T2) — is it a usable fallback for poorly filled
# owned tile sets of a sharded contraction?
mkdir -p gpurun_out/r2
for t in 128 64; do
QTB_TILE=$t timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --workload T2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('QTB_TILE=$t T2 value', round(d['value'],2), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4))
" >> gpurun_out/r2/s63.txt
done
cat gpurun_out/r2/s63.txt
