#!/bin/bash
# round-2 session 62: what one rank of a world-8 / world-4 sharded H_eff.psi runs (launch list on one GPU, collective = no-op)
mkdir -p gpurun_out/r2
for w in 8 4; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/s62_w$w.csv python profiles/r2/rank_alone.py $w > gpurun_out/r2/s62_w$w.log 2>&1
python - <<PY >> gpurun_out/r2/s62.txt
import csv
rows=list(csv.reader(open("gpurun_out/r2/s62_w$w.csv")))
hdr=None; out=[]
for r in rows:
    if len(r)>5 and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        try: v=float(d['Metric Value'].replace(',',''))
        except: continue
        out.append((d['Kernel Name'][:48], d['Grid Size'], v/1e3))
print("== world $w: last 40 launches (kernel, grid, us)")
for o in out[-40:]: print("  %-50s %-16s %9.1f" % o)
PY
done
timeout 300 python profiles/r2/rank_alone.py 8 >> gpurun_out/r2/s62.txt 2>&1
cat gpurun_out/r2/s62.txt
