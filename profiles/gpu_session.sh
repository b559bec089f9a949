#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list, ncu full capture of the grouped GEMM. Outputs -> gpurun_out/
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-extra > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 2 -c 2 -o gpurun_out/gemm_T2 \
    python profiles/prof_driver.py T2 4 > gpurun_out/ncu_T2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 2 -c 2 -o gpurun_out/gemm_T1 \
    python profiles/prof_driver.py T1 4 > gpurun_out/ncu_T1.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json
