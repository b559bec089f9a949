set -u
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
for D in 100 200 256; do
echo "== svd D=$D decay"; QTB_SVD_DEBUG=1 timeout 300 python profiles/svd_driver.py 15 $D 1.6 decay 2>&1 | tail -2
echo "== svd D=$D random"; QTB_SVD_DEBUG=1 timeout 300 python profiles/svd_driver.py 15 $D 1.6 2>&1 | tail -2
done
echo "== dmrg config0: L=50 maxbond 200 cutoff 1e-12 (23 sweeps, conv 0)"
timeout 600 python profiles/dmrg_sweep_bench.py 50 200 1e-12 23 2>&1 | tail -6
echo "== dmrg L=64 maxbond 256"
QTB_PROFILE=1 timeout 600 python profiles/dmrg_sweep_bench.py 64 256 1e-20 7 2>&1 | tail -4
