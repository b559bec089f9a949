#!/bin/bash
# round-2 session 5: column-block width / inner sweeps of the tensor-core SVD path inside a real DMRG run (D=2048)
mkdir -p gpurun_out/r2
for cfg in "16 4" "16 2" "32 2"; do
  set -- $cfg
  echo "== QTB_SVD_TCJB=$1 QTB_SVD_INNER=$2"
  QTB_SVD_TCJB=$1 QTB_SVD_INNER=$2 QTB_PROFILE=1 timeout 900 python profiles/dmrg_sweep_bench.py 64 2048 1e-20 6 2>&1 | grep -E "profile\] sweep [45]|^sweep 5"
done > gpurun_out/r2/s5.txt 2>&1
cat gpurun_out/r2/s5.txt
