set -u
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
echo "== bench default"
timeout 900 python bench.py --dmrg '' > gpurun_out/bench_T1.json 2> gpurun_out/bench_T1.err; tail -3 gpurun_out/bench_T1.err; cat gpurun_out/bench_T1.json
