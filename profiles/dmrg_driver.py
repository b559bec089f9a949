"""Runs the reference-dumped Heisenberg problem through the engine's DMRG (diagnostics / profiling driver).
usage: python profiles/dmrg_driver.py L maxbond cutoff conv maxit"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import qtb_oracle as orc
import quantit_b200 as qb

L, maxbond, cutoff, conv, maxit = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
td = tempfile.mkdtemp()
subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "heis", str(L), "4", "1e-3", "1e-1", "1", "0", td],
               check=True, capture_output=True)
eng = lambda t: qb.BTensor.from_host(t.sec_sizes, t.cvals, t.sel, t.blocks)
H = [eng(orc.read_qtbt(f"{td}/H_{i}.qtbt")) for i in range(L)]
psi = [eng(orc.read_qtbt(f"{td}/psi0_{i}.qtbt")) for i in range(L)]
log = {}
E = qb.dmrg(H, psi, qb.dmrg_options(cutoff, conv, maxbond, 4, maxit), oc=0, log=log)
print("E", E, "sweeps", len(log["energy"]), "ms/sweep", [round(1e3 * s, 1) for s in log["seconds"]], "mid", log["mid_bond"][-1])
