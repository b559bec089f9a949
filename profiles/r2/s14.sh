#!/bin/bash
# round-2 session 14 (2 GPUs): row-range sharding of the contraction chains + the engine's own NCCL exchange
mkdir -p gpurun_out/r2
( timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_configs.py -x -q 2>&1 | tail -5 ) > gpurun_out/r2/s14_pytest.txt 2>&1
for mode in "1 0" "1 1" "0 1"; do
  set -- $mode
  echo "== QTB_SHARD_MSPLIT=$1 QTB_SHARD_ALLREDUCE=$2" >> gpurun_out/r2/s14_sharded.txt
  QTB_SHARD_MSPLIT=$1 QTB_SHARD_ALLREDUCE=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/sharded_driver.py 15 4096 1.6 10 2>&1 | grep -E "world|rror|Traceback" >> gpurun_out/r2/s14_sharded.txt
done
cat gpurun_out/r2/s14_pytest.txt gpurun_out/r2/s14_sharded.txt
