#!/bin/bash
# round-2 session 15 (8 GPUs): strong scaling of one H_eff.psi at D=4096, N = 4 and 8
mkdir -p gpurun_out/r2
for n in 8 4; do
for mode in "1 0" "0 1"; do
  set -- $mode
  echo "== N=$n QTB_SHARD_MSPLIT=$1 QTB_SHARD_ALLREDUCE=$2" >> gpurun_out/r2/s15_sharded.txt
  QTB_SHARD_MSPLIT=$1 QTB_SHARD_ALLREDUCE=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 profiles/sharded_driver.py 15 4096 1.6 10 2>&1 | grep -E "world|rror|Traceback" >> gpurun_out/r2/s15_sharded.txt
done
done
cat gpurun_out/r2/s15_sharded.txt
