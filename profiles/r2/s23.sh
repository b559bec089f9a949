#!/bin/bash
mkdir -p gpurun_out/r2
QTB_SVD_DEBUG=1 SVD_REPS=3 timeout 300 python profiles/svd_driver.py 11 256 1.5 span15 2>&1 | grep -E "svd ms|lane 0" | tail -14 | cut -c1-110 > gpurun_out/r2/s23.txt
SVD_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2/s23_launches.csv python profiles/svd_driver.py 11 256 1.5 span15 > /dev/null 2>&1
python profiles/agg_launches.py gpurun_out/r2/s23_launches.csv >> gpurun_out/r2/s23.txt 2>&1
cat gpurun_out/r2/s23.txt
