#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (both arms), ncu launch lists, ncu full captures of the dominant kernels.
# Outputs -> gpurun_out/ ; the summaries that are kept go to profiles/rN/ (see profiles/README.md).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
( time timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2> gpurun_out/bench_time.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
# launch list of the bench command (the timed region of a step holds ONE kernel of ours)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-extra > gpurun_out/bench_under_ncu.log 2>&1
# launch list of H_eff.psi at D=4096 (three contractions: two DMMA launches + the HBM-bound MPO step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_heff.csv \
    python profiles/prof_driver.py HEFF 3 > gpurun_out/heff_under_ncu.log 2>&1
# launch list of one block SVD at D=4096
SVD_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_svd.csv \
    python profiles/svd_driver.py 15 4096 1.6 decay > gpurun_out/svd_under_ncu.log 2>&1
# full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 2 -c 2 -o gpurun_out/gemm_T2 \
    python profiles/prof_driver.py T2 4 > gpurun_out/ncu_T2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 2 -c 2 -o gpurun_out/gemm_T1 \
    python profiles/prof_driver.py T1 4 > gpurun_out/ncu_T1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -s 1 -c 1 -o gpurun_out/skinny_heff \
    python profiles/prof_driver.py HEFF 3 > gpurun_out/ncu_skinny.log 2>&1
SVD_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:svd_ -s 600 -c 3 -o gpurun_out/svd_step \
    python profiles/svd_driver.py 15 4096 1.6 decay > gpurun_out/ncu_svd.log 2>&1
python profiles/summarize_ncu.py gpurun_out/ncu_full_summary.csv gpurun_out/gemm_T2.ncu-rep gpurun_out/gemm_T1.ncu-rep \
    gpurun_out/skinny_heff.ncu-rep gpurun_out/svd_step.ncu-rep > gpurun_out/summarize.log 2>&1
for f in bench heff svd; do python profiles/agg_launches.py gpurun_out/launches_$f.csv > gpurun_out/launches_${f}_agg.txt 2>&1; done
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_time.txt; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json
cat gpurun_out/launches_bench_agg.txt gpurun_out/launches_heff_agg.txt; head -5 gpurun_out/launches_svd_agg.txt
