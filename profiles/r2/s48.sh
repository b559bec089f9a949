#!/bin/bash
# round-2 session 48: launch list of one D=256 panel-path SVD
mkdir -p gpurun_out/r2
SVD_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2/s48_launches.csv python profiles/svd_driver.py 11 256 1.5 span15 > /dev/null 2>&1
python profiles/agg_launches.py gpurun_out/r2/s48_launches.csv > gpurun_out/r2/s48.txt 2>&1
cat gpurun_out/r2/s48.txt
