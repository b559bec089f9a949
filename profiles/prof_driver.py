"""Tiny driver used under ncu: runs one workload's contraction a few times (no timing claims)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantit_b200 as qb
from quantit_b200 import workloads as wl

name = sys.argv[1] if len(sys.argv) > 1 else "T2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = {"T1": wl.T1, "T2": wl.T2, "T2_8K": wl.T2_8K}[name]
a, b, da, db = wl.tdot_pair(**cfg)
A, B = qb.BTensor.from_host(**a), qb.BTensor.from_host(**b)
for _ in range(reps):
    C = A.tensordot(B, da, db)
qb.default_context().sync()
print(name, A.tensordot_info(B, da, db))
