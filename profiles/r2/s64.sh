#!/bin/bash
# round-2 session 64 (8 GPUs): strong scaling of H_eff.psi with 64x64 tiles forced (QTB_TILE=64) at N = 8, 4, 2
mkdir -p gpurun_out/r2
for n in 8 4 2; do
QTB_TILE=64 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2964$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/r2/s64_n$n.json 2> gpurun_out/r2/s64_n$n.err
python - <<PY >> gpurun_out/r2/s64.txt
import json
d=json.loads(open("gpurun_out/r2/s64_n$n.json").read().strip().splitlines()[-1])
s=d["strong_scaling"]
print("QTB_TILE=64 N=$n strong: single", round(s["single_gpu_ms_per_step"],3), "sharded", round(s["ms_per_step"],3), "eff", round(s["efficiency"],3))
PY
done
cat gpurun_out/r2/s64.txt
