set -u
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
echo "== bench default"
timeout 900 python bench.py --dmrg '' > gpurun_out/bench_T1.json 2> gpurun_out/bench_T1.err; tail -3 gpurun_out/bench_T1.err; python - <<'P'
import json
d = json.load(open("gpurun_out/bench_T1.json"))
print("T1", d["value"], d["ms_per_step"], "T2", d["workloads"]["T2"]["value"], "heff", d["workloads"]["heff_D4096"]["value"], "e2e", d["e2e"]["value"])
P
echo "== T1 snake"; QTB_SCHED=snake timeout 300 python bench.py --no-extra --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
