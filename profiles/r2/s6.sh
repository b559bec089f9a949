#!/bin/bash
# round-2 session 6: launch list of one block SVD on a DMRG-like theta at D=2048 (span15 spectrum)
mkdir -p gpurun_out/r2
QTB_SVD_DEBUG=1 SVD_REPS=3 timeout 300 python profiles/svd_driver.py 15 2048 1.6 span15 2>&1 | grep -E "svd ms|lane 0" | tail -14 | cut -c1-100 > gpurun_out/r2/s6_svd.txt
SVD_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2/s6_launches.csv python profiles/svd_driver.py 15 2048 1.6 span15 > gpurun_out/r2/s6_ncu.log 2>&1
python profiles/agg_launches.py gpurun_out/r2/s6_launches.csv > gpurun_out/r2/s6_launches_agg.txt 2>&1
cat gpurun_out/r2/s6_svd.txt gpurun_out/r2/s6_launches_agg.txt
