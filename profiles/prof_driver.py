"""Tiny driver used under ncu: runs one workload's contraction a few times (no timing claims)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantit_b200 as qb
from quantit_b200 import workloads as wl

name = sys.argv[1] if len(sys.argv) > 1 else "T2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
if name == "HEFF":  # H_eff.psi at D=4096 (T3): two DMMA contractions around the HBM-bound MPO step
    psi, W, L, R = wl.heff_set(15, 4096, 1.6, seed=5)
    bt = lambda d: qb.BTensor.from_host(**d)
    Wb = bt(W)
    H2 = Wb.tensordot(Wb, [2], [0]).permute([0, 1, 3, 4, 2, 5])
    p_, l_, r_ = bt(psi), bt(L), bt(R)
    for _ in range(reps):
        out = qb.hamil2site_times_state(p_, H2, l_, r_)
    qb.default_context().sync()
    print(name, qb.default_context().counters())
    sys.exit(0)
cfg = {"T1": wl.T1, "T2": wl.T2, "T2_8K": wl.T2_8K}[name]
a, b, da, db = wl.tdot_pair(**cfg)
A, B = qb.BTensor.from_host(**a), qb.BTensor.from_host(**b)
for _ in range(reps):
    C = A.tensordot(B, da, db)
qb.default_context().sync()
print(name, A.tensordot_info(B, da, db))
