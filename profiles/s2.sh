set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_svd_dmrg.py tests/test_gpu_dmrg_config0.py -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
for D in 1024 4096; do
  echo "== svd D=$D random"; QTB_SVD_DEBUG=1 timeout 300 python profiles/svd_driver.py 15 $D 1.6 2>&1 | tail -4
done
echo "== svd D=2048 decaying"; QTB_SVD_DEBUG=1 timeout 300 python profiles/svd_driver.py 15 2048 1.6 decay 2>&1 | tail -4
echo "== svd D=2048 decaying inner 2"; QTB_SVD_INNER=2 QTB_SVD_DEBUG=1 timeout 300 python profiles/svd_driver.py 15 2048 1.6 decay 2>&1 | tail -3
echo "== svd D=2048 decaying inner 8"; QTB_SVD_INNER=8 QTB_SVD_DEBUG=1 timeout 300 python profiles/svd_driver.py 15 2048 1.6 decay 2>&1 | tail -3
echo "== dmrg L=100 maxbond 1024"
QTB_PROFILE=2 timeout 900 python profiles/dmrg_sweep_bench.py 100 1024 1e-20 6 2>&1 | tail -60
