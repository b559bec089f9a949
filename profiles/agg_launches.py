"""aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name: python profiles/agg_launches.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    agg[r[ki][:70]][0] += 1
    agg[r[ki][:70]][1] += v
tot = sum(t for _, t in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} n={n:6d} total={t/1e3:10.3f} ms ({100*t/tot:5.1f} %) avg={t/n:9.2f} us")
