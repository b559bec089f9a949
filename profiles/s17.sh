set -u
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
echo "== svd D=4096 decay"
QTB_SVD_DEBUG=1 timeout 300 python profiles/svd_driver.py 15 4096 1.6 decay 2>&1 | tail -2
echo "== svd D=4096 random"
timeout 300 python profiles/svd_driver.py 15 4096 1.6 2>&1 | tail -1
echo "== svd D=1024 decay"
timeout 300 python profiles/svd_driver.py 15 1024 1.6 decay 2>&1 | tail -1
echo "== ncu launch list of one SVD"
SVD_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/svd_launches.csv python profiles/svd_driver.py 15 4096 1.6 decay > gpurun_out/svd_ncu.log 2>&1
python profiles/agg_launches.py gpurun_out/svd_launches.csv > gpurun_out/svd_agg.txt; head -4 gpurun_out/svd_agg.txt
echo "== dmrg L=100 maxbond 4096 (QTB_PROFILE=1: per-phase wall time)"
QTB_PROFILE=1 timeout 1500 python profiles/dmrg_sweep_bench.py 100 4096 1e-20 7 2>&1 | tail -18
