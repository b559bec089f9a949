set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== sharded driver N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/sharded_driver.py 15 4096 1.6 5 2>&1 | grep -v Warning | tail -15
echo "== bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
echo "== bench N=2 reference arm"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -2
