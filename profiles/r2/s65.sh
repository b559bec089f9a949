#!/bin/bash
# round-2 session 65 (2 GPUs): 64x64 tiles chosen automatically for sharded plans: sharded tests, bit-identity at D=4096, bench N=2
mkdir -p gpurun_out/r2
( timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_tensordot.py -x -q 2>&1 | tail -3 ) > gpurun_out/r2/s65.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 profiles/sharded_driver.py 15 4096 1.6 10 2>&1 | grep -E "world|rror|Traceback" >> gpurun_out/r2/s65.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2/s65_n2.json 2> gpurun_out/r2/s65_n2.err
python - <<PY >> gpurun_out/r2/s65.txt
import json
d=json.loads(open("gpurun_out/r2/s65_n2.json").read().strip().splitlines()[-1])
s=d["strong_scaling"]
print("N=2 value", round(d["value"],2), "strong: single", round(s["single_gpu_ms_per_step"],3), "sharded", round(s["ms_per_step"],3), "eff", round(s["efficiency"],3))
PY
cat gpurun_out/r2/s65.txt
