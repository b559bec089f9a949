#!/bin/bash
# Round-2 evidence run (one gpurun call, one B200): smoke, bench (both arms), ncu launch lists of the bench command,
# of H_eff.psi at D=4096 and of block SVDs on both Jacobi paths, ncu --set full captures of the dominant kernels.
# Outputs -> gpurun_out/r2final/ ; what is kept is copied to profiles/r2/ (profiles/README.md).
set -u
O=gpurun_out/r2final
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
( time timeout 1200 python bench.py > $O/bench_T1.json 2> $O/bench.err ) 2> $O/bench_time.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_T1_reference.json 2>> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_T1.csv \
    python bench.py --steps 5 --warmup 3 --no-extra > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_heff_D4096.csv \
    python profiles/prof_driver.py HEFF 3 > $O/heff_under_ncu.log 2>&1
SVD_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_svd_D4096.csv \
    python profiles/svd_driver.py 31 4096 3.2 span15 > $O/svd4096_under_ncu.log 2>&1
SVD_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_svd_D256.csv \
    python profiles/svd_driver.py 11 256 1.5 span15 > $O/svd256_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 2 -c 2 -o $O/gemm_T2 \
    python profiles/prof_driver.py T2 4 > $O/ncu_T2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 2 -c 2 -o $O/gemm_T1 \
    python profiles/prof_driver.py T1 4 > $O/ncu_T1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -s 1 -c 1 -o $O/skinny_heff \
    python profiles/prof_driver.py HEFF 3 > $O/ncu_skinny.log 2>&1
SVD_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:svd_fused -s 300 -c 2 -o $O/svd_fused \
    python profiles/svd_driver.py 31 4096 3.2 span15 > $O/ncu_svd_fused.log 2>&1
SVD_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:svd_panel -s 60 -c 2 -o $O/svd_panel \
    python profiles/svd_driver.py 11 256 1.5 span15 > $O/ncu_svd_panel.log 2>&1
SVD_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:qr_panel -s 4 -c 1 -o $O/qr_panel \
    python profiles/svd_driver.py 31 4096 3.2 span15 > $O/ncu_qr_panel.log 2>&1
python profiles/summarize_ncu.py $O/ncu_full_summary.csv $O/gemm_T2.ncu-rep $O/gemm_T1.ncu-rep $O/skinny_heff.ncu-rep \
    $O/svd_fused.ncu-rep $O/svd_panel.ncu-rep $O/qr_panel.ncu-rep > $O/summarize.log 2>&1
for f in bench_T1 heff_D4096 svd_D4096 svd_D256; do python profiles/agg_launches.py $O/launches_$f.csv > $O/launches_${f}_agg.txt 2>&1; done
rm -f $O/*.ncu-rep
tail -2 $O/smoke.log; cat $O/bench_time.txt; cat $O/bench_T1_reference.json; head -4 $O/launches_bench_T1_agg.txt; head -4 $O/launches_heff_D4096_agg.txt
