set -u
mkdir -p gpurun_out
echo "== T1 lpt"; timeout 300 python bench.py --no-extra --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
echo "== T1 snake"; QTB_SCHED=snake timeout 300 python bench.py --no-extra --steps 300 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -s 2 -c 1 -o gpurun_out/gemm_T1 \
    python profiles/prof_driver.py T1 4 > gpurun_out/ncu_T1.log 2>&1
ls -la gpurun_out/*.ncu-rep
