#!/bin/bash
# round-2 session 59: T4 at the size BASELINE.json configs[4] names: truncated SVD of a Hubbard U(1)xU(1) theta at D = 8192
mkdir -p gpurun_out/r2
timeout 900 python - > gpurun_out/r2/s59.txt 2>&1 <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import bench
import quantit_b200 as qb
ctx = qb.default_context()
print(json.dumps(bench.time_svd_sweep(torch, qb, ctx, [8192], ref_max_D=0)))
PY
cat gpurun_out/r2/s59.txt | tail -5
