#!/bin/bash
# round-2 session 2: bulk-copy (UBLKCP) operand staging in the grouped GEMM: parity, then T1 / T2 / H_eff timings with and without
mkdir -p gpurun_out/r2
( timeout 600 python -m pytest tests/test_gpu_tensordot.py -x -q 2>&1 | tail -5 ) > gpurun_out/r2/s2_pytest.txt
for bulk in 1 0; do
  echo "== QTB_BULK=$bulk"
  QTB_BULK=$bulk timeout 300 python bench.py --steps 100 --warmup 5 --dmrg '' 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('T1 value', d['value'], d['unit'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
for k, v in d.get('workloads', {}).items(): print(' ', k, v)
"
done > gpurun_out/r2/s2_bench.txt 2>&1
cat gpurun_out/r2/s2_pytest.txt gpurun_out/r2/s2_bench.txt
