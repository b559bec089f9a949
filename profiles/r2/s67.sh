#!/bin/bash
# round-2 session 67: final build: whole GPU suite, smoke, default bench
O=gpurun_out/r2final3
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > $O/pytest_gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 > $O/smoke.log
( time timeout 1200 python bench.py > $O/bench_T1.json 2> $O/bench.err ) 2> $O/bench_time.txt
cat $O/pytest_gpu.txt $O/smoke.log $O/bench_time.txt; tail -3 $O/bench.err
python - <<PY
import json
d=json.loads(open("$O/bench_T1.json").read().strip().splitlines()[-1])
w=d["workloads"]
print("T1", round(d["ms_per_step"]*1e3,1), "us frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],3))
print("T2", round(w["T2"]["value"],2), "heff", round(w["heff_D4096"]["value"],2), "env", w["env_left_D4096"])
print("svd", {k:round(v["ms"],1) for k,v in w["svd_sweep_hubbard"].items()})
for k in ("dmrg_sweep_L64_D256","dmrg_sweep_L100_D4096"): print(k, [round(x,2) for x in w[k]["sweep_seconds"]])
PY
