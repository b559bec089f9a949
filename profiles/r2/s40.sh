#!/bin/bash
# round-2 session 40: which side bounds the T1 kernel: run without DMMAs (1), without copies (2), without both (3)
mkdir -p gpurun_out/r2
for v in 0 1 2 3; do
  QTB_GEMM_DEBUG=$v timeout 300 python bench.py --steps 100 --warmup 5 --no-extra --workload T1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('QTB_GEMM_DEBUG=$v T1 ms', round(d['ms_per_step'],5))
" >> gpurun_out/r2/s40.txt
done
cat gpurun_out/r2/s40.txt
