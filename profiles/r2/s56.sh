#!/bin/bash
# round-2 session 56: final build (column pre-sort included): whole GPU suite, smoke, default bench, reference arm,
# launch lists of the two SVD paths
O=gpurun_out/r2final2
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > $O/pytest_gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 > $O/smoke.log
( time timeout 1200 python bench.py > $O/bench_T1.json 2> $O/bench.err ) 2> $O/bench_time.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_T1_reference.json 2>> $O/bench.err
SVD_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_svd_D4096.csv \
    python profiles/svd_driver.py 31 4096 3.2 span15 > /dev/null 2>&1
SVD_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_svd_D256.csv \
    python profiles/svd_driver.py 11 256 1.5 span15 > /dev/null 2>&1
for f in svd_D4096 svd_D256; do python profiles/agg_launches.py $O/launches_$f.csv > $O/launches_${f}_agg.txt 2>&1; done
QTB_SVD_DEBUG=2 timeout 600 python profiles/dmrg_sweep_bench.py 100 4096 1e-20 6 2>&1 | grep -E "phases ms|census|lane 0|^sweep" | awk '/census/{print prev; print} /phases ms/{print} {prev=$0} /^sweep/{print}' | tail -130 | cut -c1-170 > $O/dmrg4096_svd_calls.txt
cat $O/pytest_gpu.txt $O/smoke.log $O/bench_time.txt; head -3 $O/launches_svd_D4096_agg.txt | cut -c1-130; head -2 $O/launches_svd_D256_agg.txt | cut -c1-130; tail -12 $O/dmrg4096_svd_calls.txt
