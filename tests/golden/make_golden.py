"""Generates tests/golden/*.qtbt by running the UNMODIFIED reference (oracle/_ref/ref_harness, built by
`make -C oracle ref` from /root/reference) on seeded inputs. Run in the build container only:

    python tests/golden/make_golden.py

Inputs come from quantit_b200.workloads (seeded numpy); outputs are what the reference's own C++ code computed:
btensor::tensordot, svd(btensor,split[,tol,min,max,pow]), hamil2site_times_state, compute_left/right_env,
two_sites_update. The fixtures pin oracle/qtb_oracle.py (tests/test_oracle_vs_reference.py) and, on the GPU box where
/root/reference does not exist, the CUDA engine (tests/test_gpu_golden.py).
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import qtb_oracle as orc  # noqa: E402
from quantit_b200 import workloads as wl  # noqa: E402

H = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def bt(d):
    return orc.BT(**{k: d[k] for k in ("sec_sizes", "cvals", "sel", "blocks")})


def w(name, t):
    orc.write_qtbt(t, os.path.join(HERE, name + ".qtbt"))
    return os.path.join(HERE, name + ".qtbt")


def run(*args):
    subprocess.check_call([H] + [str(a) for a in args])


def main():
    rng = np.random.default_rng(2024)
    # ---- tensordot ----------------------------------------------------------------------------------------------
    a, b, da, db = wl.tdot_pair(7, 40, 1.3, seed=21)
    run("tdot", w("tdot1_A", bt(a)), w("tdot1_B", bt(b)), "3", "0", os.path.join(HERE, "tdot1_C.qtbt"))
    # multi-index contraction in scrambled order, two-component charges, missing blocks
    a2 = wl.random_btensor(rng, 4, max_sec=3, max_size=4, nc=2, fill=0.8)
    legs = [None] * 3
    legs[2] = wl.conj_leg((a2["sec_sizes"][1], a2["cvals"][1]))
    legs[0] = wl.conj_leg((a2["sec_sizes"][3], a2["cvals"][3]))
    legs[1] = ([2, 3], [(1, 0), (-1, 1)])
    b2 = wl.random_btensor(rng, 3, nc=2, legs=legs, fill=0.8)
    run("tdot", w("tdot2_A", bt(a2)), w("tdot2_B", bt(b2)), "1,3", "2,0", os.path.join(HERE, "tdot2_C.qtbt"))
    # outer product (no contracted dims) and full trace (rank-0 result)
    a3 = wl.random_btensor(rng, 2, max_sec=3, fill=1.0)
    b3 = wl.random_btensor(rng, 2, max_sec=3, fill=1.0)
    run("tdot", w("tdot3_A", bt(a3)), w("tdot3_B", bt(b3)), "-", "-", os.path.join(HERE, "tdot3_C.qtbt"))
    a4 = bt(wl.random_btensor(rng, 3, max_sec=3, fill=1.0))
    run("tdot", w("tdot4_A", a4), w("tdot4_B", orc.conj(a4)), "2,0,1", "2,0,1", os.path.join(HERE, "tdot4_C.qtbt"))
    # ---- H_eff, environments, Lanczos update --------------------------------------------------------------------
    psi, W, L, R = wl.heff_set(5, 20, 1.1, seed=5)
    oW = bt(W)
    oH2 = orc.permute(orc.tensordot(oW, oW, [2], [0]), [0, 1, 3, 4, 2, 5])
    f = {n: w("heff_" + n, t) for n, t in [("psi", bt(psi)), ("H2", oH2), ("L", bt(L)), ("R", bt(R)), ("W", oW)]}
    run("heff", f["psi"], f["H2"], f["L"], f["R"], os.path.join(HERE, "heff_out.qtbt"))
    run("update", f["psi"], f["H2"], f["L"], f["R"], os.path.join(HERE, "update_E.qtbt"),
        os.path.join(HERE, "update_psi.qtbt"))
    # an MPS site Y[a,s,b] compatible with L / R and W for the env updates
    beta = wl.bond(5, 20, 1.1, 2)
    Y = wl.rand_like(wl.shape([beta, wl.SPIN_HALF, wl.conj_leg(beta)], (-1,)), np.random.default_rng(6))
    fy = w("env_Y", bt(Y))
    run("lenv", f["W"], fy, f["L"], os.path.join(HERE, "env_left.qtbt"))
    Y2 = wl.rand_like(wl.shape([beta, wl.SPIN_HALF, wl.conj_leg(beta)], (1,)), np.random.default_rng(7))
    fy2 = w("env_Y2", bt(Y2))
    run("renv", f["W"], fy2, f["R"], os.path.join(HERE, "env_right.qtbt"))
    # ---- block SVD with and without truncation ------------------------------------------------------------------
    th = wl.rand_like(wl.shape([wl.bond(5, 24, 1.2, 2), wl.SPIN_HALF, wl.SPIN_HALF, wl.conj_leg(wl.bond(5, 30, 1.2, 2))],
                               (0,)), np.random.default_rng(8))
    # make it rapidly decaying so that truncation has something to do
    for k in th["blocks"]:
        u, s, vt = np.linalg.svd(th["blocks"][k].reshape(th["blocks"][k].shape[0], -1), full_matrices=False)
        s = s * np.exp(-1.5 * np.arange(len(s)))
        th["blocks"][k] = ((u * s) @ vt).reshape(th["blocks"][k].shape)
    ft = w("svd_theta", bt(th))
    run("svd", ft, 2, *(os.path.join(HERE, f"svd_{x}.qtbt") for x in "UdV"))
    run("svdt", ft, 2, 1e-4, 4, 18, 2, *(os.path.join(HERE, f"svdt_{x}.qtbt") for x in "UdV"))
    run("svdt", ft, 2, 1e-1, 1, 1000, 2, *(os.path.join(HERE, f"svdt2_{x}.qtbt") for x in "UdV"))
    run("svdt", ft, 2, 0.5, 1, 1000, 2, *(os.path.join(HERE, f"svdt3_{x}.qtbt") for x in "UdV"))  # drops whole sectors
    # ---- whole two-site DMRG: MPO + initial MPS dumped by the reference, its per-sweep energies, its final MPS ----
    import json
    for name, cmd, L, maxbond in [("dmrg_heis8", "heis", 8, 30), ("dmrg_hub4", "hub", 4, 40)]:
        d = os.path.join(HERE, name)
        os.makedirs(d, exist_ok=True)
        out = subprocess.run([H, cmd, str(L), str(maxbond), "1e-10", "1e-9", "12", "0", d], capture_output=True, text=True,
                             check=True).stdout
        rec = {"L": L, "maximum_bond": maxbond, "cutoff": 1e-10, "convergence_criterion": 1e-9, "minimum_bond": 4,
               "maximum_iterations": 12,
               "sweep_energy": [float(l.split()[3]) for l in out.splitlines() if l.startswith("SWEEP")],
               "mid_bond": [int(l.split()[5]) for l in out.splitlines() if l.startswith("SWEEP")],
               "E0": [float(l.split()[1]) for l in out.splitlines() if l.startswith("E0")][0],
               "contract_E": [float(l.split()[1]) for l in out.splitlines() if l.startswith("CONTRACT_E")][0],
               "oc": [int(l.split()[1]) for l in out.splitlines() if l.startswith("OC")][0]}
        rec["norm"] = [float(l.split()[3]) for l in out.splitlines() if l.startswith("CONTRACT_E")][0]
        json.dump(rec, open(os.path.join(d, "reference_run.json"), "w"), indent=1)
        # ---- bMPS::move_oc on the final MPS: centre 0 -> L-2 (moves right), then L-2 -> 1 (moves left) ----
        for tag, prefix, oc, target in [("psiM", "psiF", rec["oc"], L - 2), ("psiN", "psiM", L - 2, 1)]:
            td = os.path.join(d, "_tmp")
            os.makedirs(td, exist_ok=True)
            subprocess.run([H, "moveoc", d, prefix, str(L), str(oc), str(target), td, d], check=True, capture_output=True)
            for i in range(L):
                os.replace(os.path.join(td, f"psiM_{i}.qtbt"), os.path.join(td, f"{tag}x_{i}.qtbt"))
            for i in range(L):
                os.replace(os.path.join(td, f"{tag}x_{i}.qtbt"), os.path.join(d, f"{tag}_{i}.qtbt"))
            os.rmdir(td)
        # ---- bMPO::coalesce(1e-10) of the MPO: Hc_i.qtbt ----
        subprocess.run([H, "coalesce", d, str(L), "1e-10", d], check=True, capture_output=True)
    tot = sum(os.path.getsize(os.path.join(HERE, x)) for x in os.listdir(HERE) if x.endswith(".qtbt"))
    print("golden fixtures written,", tot, "bytes")


if __name__ == "__main__":
    main()
