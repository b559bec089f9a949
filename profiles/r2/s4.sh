#!/bin/bash
# round-2 session 4: what do the theta' spectra of a real sweep look like, and what does QR buy inside a DMRG run
mkdir -p gpurun_out/r2
for qr in 1 0; do
  echo "== QTB_SVD_QR=$qr"
  QTB_SVD_QR=$qr QTB_SVD_DEBUG=2 QTB_PROFILE=1 timeout 900 python profiles/dmrg_sweep_bench.py 64 2048 1e-20 6 2>&1 | grep -E "census|sweep [0-9]+ lane 0|profile\] sweep|^sweep" > gpurun_out/r2/s4_dmrg_qr$qr.txt
  grep -E "profile\] sweep|^sweep" gpurun_out/r2/s4_dmrg_qr$qr.txt
done
grep census gpurun_out/r2/s4_dmrg_qr1.txt | tail -130 | awk 'NR%8==0' | cut -c1-200
