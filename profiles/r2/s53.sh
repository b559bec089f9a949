#!/bin/bash
# round-2 session 53 (8 GPUs): bench.py under torchrun at N = 8 and 4 with the final build (weak scaling of the headline +
# strong scaling of ONE H_eff.psi at D=4096)
mkdir -p gpurun_out/r2
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/r2/bench_T1_n$n.json 2> gpurun_out/r2/s53_n$n.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2/bench_T1_n$n.json").read().strip().splitlines()[-1])
s=d["strong_scaling"]
print("N=$n value", round(d["value"],2), "ms", round(d["ms_per_step"],5), "strong: single", round(s["single_gpu_ms_per_step"],3), "sharded", round(s["ms_per_step"],3), "eff", round(s["efficiency"],3), s["exchange"][:20])
PY
done
