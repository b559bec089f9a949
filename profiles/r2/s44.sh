#!/bin/bash
# round-2 session 44: BK=32, 3 stages on the 64x64 configuration: tests + T1/T2 + skeleton diagnostics
mkdir -p gpurun_out/r2
( timeout 1200 python -m pytest tests -m gpu -x -q -k "tensordot or gemm or contract or config or heff or capi or dmrg" 2>&1 | tail -3 ) > gpurun_out/r2/s44.txt
for w in T1 T2; do
  timeout 300 python bench.py --steps 100 --warmup 5 --no-extra --workload $w 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w value', round(d['value'],3), d['unit'], 'ms', round(d['ms_per_step'],5), 'frac', round(d['roofline']['frac'],4))
" >> gpurun_out/r2/s44.txt
done
for v in 1 2 3; do
  QTB_GEMM_DEBUG=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-extra --workload T1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('QTB_GEMM_DEBUG=$v T1 ms', round(d['ms_per_step'],5))
" >> gpurun_out/r2/s44.txt
done
cat gpurun_out/r2/s44.txt
