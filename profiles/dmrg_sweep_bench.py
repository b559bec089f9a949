"""DMRG sweep time at a saturated bond dimension (diagnostics driver; bench.py's "dmrg_sweep" workload is the judged one).
usage: python profiles/dmrg_sweep_bench.py L maxbond cutoff n_sweeps
Inputs (Heisenberg U(1) bMPO + random bond-4 bMPS) are dumped by the compiled reference (oracle/_ref/ref_harness heis);
the sweeps run with convergence_criterion = 0 so that exactly n_sweeps sweeps are timed."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qtb_oracle as orc
import quantit_b200 as qb

L, maxbond, cutoff, nsw = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
td = tempfile.mkdtemp()
subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "heis", str(L), "4", "1e-3", "1e-1", "1", "0", td],
               check=True, capture_output=True)
eng = lambda t: qb.BTensor.from_host(t.sec_sizes, t.cvals, t.sel, t.blocks)
H = [eng(orc.read_qtbt(f"{td}/H_{i}.qtbt")) for i in range(L)]
psi = [eng(orc.read_qtbt(f"{td}/psi0_{i}.qtbt")) for i in range(L)]
log = {}
ctx = qb.default_context()
c0 = ctx.counters()
E = qb.dmrg(H, psi, qb.dmrg_options(cutoff, 0.0, maxbond, 4, nsw), oc=0, log=log)
c1 = ctx.counters()
for i in range(len(log["energy"])):
    print(f"sweep {i}: E {log['energy'][i]:.12f} mid_bond {log['mid_bond'][i]} seconds {log['seconds'][i]:.3f}", flush=True)
print("gemm_flops total", c1["gemm_flops"] - c0["gemm_flops"], "launches", c1["kernel_launches"] - c0["kernel_launches"],
      "device_bytes", c1["device_bytes"])
