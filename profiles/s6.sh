set -u
mkdir -p gpurun_out
echo "== svd D=4096 decay (plain)"
QTB_SVD_DEBUG=1 timeout 300 python profiles/svd_driver.py 15 4096 1.6 decay 2>&1 | tail -4
echo "== ncu launch list of one SVD"
SVD_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/svd_launches.csv python profiles/svd_driver.py 15 4096 1.6 decay > gpurun_out/svd_ncu.log 2>&1
tail -2 gpurun_out/svd_ncu.log
python - <<'P'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/svd_launches.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else v
    agg[r[ki][:60]][0] += 1; agg[r[ki][:60]][1] += v
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} n={n:6d} total={t/1e3:10.3f} ms avg={t/n:9.2f} us")
P
