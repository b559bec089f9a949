"""times the block SVD of a DMRG-like two-site tensor (diagnostics): python profiles/svd_driver.py n_sec D sigma"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quantit_b200 as qb
from quantit_b200 import workloads as wl
n_sec, D, sigma = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
rng = np.random.default_rng(1)
beta = wl.bond(n_sec, D, sigma, 2)
th = wl.rand_like(wl.shape([beta, wl.SPIN_HALF, wl.SPIN_HALF, wl.conj_leg(beta)], (0,)), rng)
if len(sys.argv) > 4:  # decaying spectrum like a real DMRG theta
    for k in th["blocks"]:
        b = th["blocks"][k]; u, s, vt = np.linalg.svd(b.reshape(b.shape[0], -1), full_matrices=False)
        # "decay": exp(-0.15 k) (numerically rank deficient); "span15": log-uniform over 15 decades per block, the profile
        # of the theta' of a real D=2048 sweep (profiles/r2/s4_dmrg_qr1.txt census lines)
        prof = np.exp(-0.15 * np.arange(len(s))) if sys.argv[4] == "decay" else 10.0 ** (-15.0 * np.arange(len(s)) / max(len(s) - 1, 1))
        th["blocks"][k] = ((u * (s[0] * prof)) @ vt).reshape(b.shape)
T = qb.BTensor.from_host(**th)
ctx = qb.default_context()
for i in range(int(os.environ.get("SVD_REPS", "3"))):
    ctx.sync(); t0 = time.perf_counter()
    U, d, V = qb.svd(T, 2, 1e-12, 4, D)
    ctx.sync(); print("svd ms", (time.perf_counter() - t0) * 1e3, "kept", sum(d.structure()[0][0]))
