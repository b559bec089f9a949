mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2/s25_pytest.txt
cat gpurun_out/r2/s25_pytest.txt
QTB_PROFILE=1 timeout 300 python profiles/dmrg_sweep_bench.py 100 256 1e-20 6 2>&1 | grep -E "profile\] sweep|^sweep" > gpurun_out/r2/s25_dmrg256.txt; tail -4 gpurun_out/r2/s25_dmrg256.txt
