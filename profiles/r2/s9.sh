#!/bin/bash
# round-2 session 9: ncu of the fused SVD kernel (launch durations + one full capture)
mkdir -p gpurun_out/r2
SVD_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 1500 --csv --log-file gpurun_out/r2/s9_launches.csv python profiles/svd_driver.py 15 2048 1.6 span15 > gpurun_out/r2/s9_ncu.log 2>&1
python profiles/agg_launches.py gpurun_out/r2/s9_launches.csv > gpurun_out/r2/s9_launches_agg.txt 2>&1
SVD_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:svd_fused -s 300 -c 2 -o gpurun_out/r2/s9_fused python profiles/svd_driver.py 15 2048 1.6 span15 > gpurun_out/r2/s9_ncu2.log 2>&1
cat gpurun_out/r2/s9_launches_agg.txt; ls -la gpurun_out/r2/
