set -u
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
echo "== dmrg L=100 maxbond 1024"
QTB_PROFILE=2 timeout 600 python profiles/dmrg_sweep_bench.py 100 1024 1e-20 6 2>&1 | tail -16
echo "== dmrg L=100 maxbond 4096"
QTB_PROFILE=1 timeout 1200 python profiles/dmrg_sweep_bench.py 100 4096 1e-20 8 2>&1 | tail -20
