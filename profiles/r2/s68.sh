#!/bin/bash
# round-2 session 68 (4 GPUs): bench.py at N = 4 with the final build
mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29684 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/r2/bench_T1_n4.json 2> gpurun_out/r2/s68.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2/bench_T1_n4.json").read().strip().splitlines()[-1])
s=d["strong_scaling"]
print("N=4 value", round(d["value"],2), "strong: single", round(s["single_gpu_ms_per_step"],3), "sharded", round(s["ms_per_step"],3), "eff", round(s["efficiency"],3))
PY
