#!/bin/bash
# round-2 session 52 (2 GPUs): bench.py both arms under torchrun, N=2, with the final build
mkdir -p gpurun_out/r2
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 ) > gpurun_out/r2/bench_T1_n2.json 2> gpurun_out/r2/s52.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2/bench_T1_n2_reference.json 2>> gpurun_out/r2/s52.err
tail -4 gpurun_out/r2/s52.err; cat gpurun_out/r2/bench_T1_n2.json; cat gpurun_out/r2/bench_T1_n2_reference.json | cut -c1-300
