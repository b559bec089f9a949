#!/bin/bash
# round-2 session 16 (8 GPUs): strong scaling with the staged all-gather exchange, N = 8, 4, 2
mkdir -p gpurun_out/r2
( timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -3 ) > gpurun_out/r2/s16_sharded.txt 2>&1
for n in 8 4 2; do
  echo "== N=$n" >> gpurun_out/r2/s16_sharded.txt
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 profiles/sharded_driver.py 15 4096 1.6 10 2>&1 | grep -E "world|rror|Traceback" >> gpurun_out/r2/s16_sharded.txt
done
cat gpurun_out/r2/s16_sharded.txt
