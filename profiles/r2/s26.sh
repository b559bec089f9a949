#!/bin/bash
# round-2 session 26: where the D=4096 SVD time goes (phase cycles of the fused kernel + launch list)
mkdir -p gpurun_out/r2
QTB_SVD_DEBUG=3 SVD_REPS=2 timeout 300 python profiles/svd_driver.py 31 4096 3.2 span15 2>&1 | grep -vE "lane [1-9]" | cut -c1-160 | tail -40 > gpurun_out/r2/s26.txt
SVD_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2/s26_launches.csv python profiles/svd_driver.py 31 4096 3.2 span15 > /dev/null 2>&1
python profiles/agg_launches.py gpurun_out/r2/s26_launches.csv >> gpurun_out/r2/s26.txt 2>&1
cat gpurun_out/r2/s26.txt
