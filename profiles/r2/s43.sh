#!/bin/bash
# round-2 session 43: T1 kernel pieces: 4 = no epilogue stores, 7 = handshakes + descriptors only, 5 / 6
mkdir -p gpurun_out/r2
for v in 0 4 5 6 7; do
  QTB_GEMM_DEBUG=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-extra --workload T1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('QTB_GEMM_DEBUG=$v T1 ms', round(d['ms_per_step'],5))
" >> gpurun_out/r2/s43.txt
done
cat gpurun_out/r2/s43.txt
