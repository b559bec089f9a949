"""GPU parity tests on the BASELINE.json configurations beyond configs[0]:

* the reference's own 4x8 Heisenberg cylinder fixture (tests/2Dheisenberg.cpp:318-330) — the only input whose MPO bond
  sections are wider than 4 (5..10 here): committed golden inputs + reference sweep energies (tests/golden/cyl, made by
  tests/golden/make_cyl_golden.sh), and a live run of the compiled reference at its own default options;
* Fermi-Hubbard U(1)xU(1), L=10, bond 100 (E = -25.3806188);
* per-update parity at LARGE bond inside a real run: the two-site tensor of a Heisenberg L=50 state at bond 450 (rows
  + cols = 1800 > 879: the tensor-core / fused SVD path with QR preconditioning), its truncated SVD and its Lanczos
  update against the compiled reference on the same dumped inputs.

Tolerances: sweep 0 is the same floating-point problem on both sides: 1e-10 relative (north star). Later sweeps
separate as soon as one singular value sits on the truncation threshold (gesdd vs Jacobi last bits); converged energies
are compared at the level the stopping criterion allows (stated at each assert)."""
import os
import subprocess

import numpy as np
import pytest

import qtb_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
CYL = os.path.join(ROOT, "oracle", "_ref", "ref_cyl")
GOLD = os.path.join(ROOT, "tests", "golden", "cyl")
THREADS = str(min(16, os.cpu_count() or 1))


def eng(qb, t: orc.BT):
    return qb.BTensor.from_host(t.sec_sizes, t.cvals, t.sel, t.blocks, mods=t.mods)


def back(t) -> orc.BT:
    ss, cv, sel, mods = t.structure()
    return orc.BT(ss, cv, sel, t.to_host(), None if not any(mods) else mods)


def sweeps_of(text):
    rows = [l.split() for l in text.splitlines() if l.startswith("SWEEP")]
    return [float(r[3]) for r in rows], [int(r[5]) for r in rows]


def load_chain(qb, d, L, prefix):
    return [eng(qb, orc.read_qtbt(os.path.join(d, f"{prefix}_{i}.qtbt"))) for i in range(L)]


def test_cylinder_fixture_golden(engine):
    """committed inputs of the reference's 4x8 cylinder + its sweep energies at maximum_bond 48"""
    qb = engine
    ref_E, ref_mid = sweeps_of(open(os.path.join(GOLD, "reference_sweeps.txt")).read())
    H = load_chain(qb, GOLD, 32, "H")
    psi = load_chain(qb, GOLD, 32, "psi0")
    widest = max(max(h.structure()[0][2]) for h in H)
    assert widest >= 8, "the fixture is there for its wide MPO bond sections"
    log = {}
    E = qb.dmrg(H, psi, qb.dmrg_options(1e-12, 0.0, 48, 4, 5), oc=0, log=log)
    print("\n   reference sweeps:", " ".join(f"{e:.10f}" for e in ref_E))
    print("   engine    sweeps:", " ".join(f"{e:.10f}" for e in log["energy"]))
    assert abs(log["energy"][0] - ref_E[0]) <= 1e-10 * abs(ref_E[0])
    # the trajectories stay together to the truncation-tie level while the bond is still growing, then drift
    assert abs(log["energy"][1] - ref_E[1]) <= 1e-7 * abs(ref_E[1])
    assert abs(E - ref_E[-1]) <= 2e-3 * abs(ref_E[-1])
    assert abs(log["mid_bond"][-1] - ref_mid[-1]) <= 3
    # self-consistency: <psi|H|psi> / <psi|psi> of the final state is the Lanczos energy of the last update
    assert abs(qb.contract(psi, psi, H) / qb.contract(psi, psi) - E) <= 1e-6 * abs(E)


@pytest.mark.skipif(not os.path.exists(CYL), reason="compiled reference (oracle/_ref/ref_cyl) not present")
def test_cylinder_live_reference_default_options(engine, tmp_path):
    """the reference's own test settings (default cutoff 1e-6, convergence 1e-5) at maximum_bond 300: both runs stop
    within the stopping criterion of the same energy"""
    qb = engine
    out = subprocess.run([CYL, str(tmp_path), "300", "1e-6", "1e-5", "50", "0", "--threads", THREADS],
                         capture_output=True, text=True, check=True).stdout
    ref_E, ref_mid = sweeps_of(out)
    H = load_chain(qb, str(tmp_path), 32, "H")
    psi = load_chain(qb, str(tmp_path), 32, "psi0")
    log = {}
    E = qb.dmrg(H, psi, qb.dmrg_options(1e-6, 1e-5, 300, 4, 50), oc=0, log=log)
    print(f"\n4x8 cylinder, bond <= 300: reference E0={ref_E[-1]:.10f} in {len(ref_E)} sweeps; engine E0={E:.10f} in "
          f"{len(log['energy'])} sweeps; mid bond {log['mid_bond'][-1]} vs {ref_mid[-1]}")
    assert abs(log["energy"][0] - ref_E[0]) <= 1e-10 * abs(ref_E[0])
    assert abs(E - ref_E[-1]) <= 2e-5 * abs(ref_E[-1])  # convergence_criterion 1e-5 is what both runs stop on


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="compiled reference (oracle/_ref) not present")
def test_hubbard_L10_bond100(engine, tmp_path):
    """Fermi-Hubbard chain, U(1)xU(1) (charge, Sz), U=4, mu=2, N=L=10, Sz=0"""
    qb = engine
    out = subprocess.run([HARNESS, "hub", "10", "100", "1e-12", "1e-10", "60", "0", str(tmp_path), "--threads", THREADS],
                         capture_output=True, text=True, check=True).stdout
    ref_E, ref_mid = sweeps_of(out)
    H = load_chain(qb, str(tmp_path), 10, "H")
    psi = load_chain(qb, str(tmp_path), 10, "psi0")
    log = {}
    E = qb.dmrg(H, psi, qb.dmrg_options(1e-12, 1e-10, 100, 4, 60), oc=0, log=log)
    print(f"\nHubbard L=10 bond<=100: reference E0={ref_E[-1]:.12f} ({len(ref_E)} sweeps), engine E0={E:.12f} "
          f"({len(log['energy'])} sweeps)")
    # Sweep 0 is NOT bit-comparable here: the (charge, Sz) multiplets make exactly degenerate singular values, a tie on
    # the truncation threshold is resolved by the last bits of gesdd vs Jacobi, and with minimum_bond = 4 on a random
    # bond-4 start that already happens in the first sweep (observed 1.6e-4). The converged energy is the parity statement.
    assert abs(log["energy"][0] - ref_E[0]) <= 1e-3 * abs(ref_E[0])
    # BASELINE.md quotes -25.380618801436 (19 sweeps, 1 thread); the reference itself lands on -25.3806188082 with 8
    # threads (22 sweeps): runs that stop on a 1e-10 relative change agree to ~1e-9
    assert abs(E - ref_E[-1]) <= 2e-9 * abs(ref_E[-1])
    assert abs(E - (-25.380618801436)) <= 2e-9 * 25.38


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="compiled reference (oracle/_ref) not present")
def test_large_bond_update_and_svd_inside_a_real_run(engine, tmp_path):
    """Heisenberg L=50 after 4 reference sweeps at maximum_bond 450: the centre two-site tensor of THAT state (rows + cols
    = 1800: the tensor-core SVD path), its truncated SVD and one Lanczos update, engine vs compiled reference"""
    qb = engine
    L, D = 50, 450
    subprocess.run([HARNESS, "heis", str(L), str(D), "1e-12", "0.0", "4", "0", str(tmp_path), "--threads", THREADS],
                   capture_output=True, text=True, check=True)
    H = load_chain(qb, str(tmp_path), L, "H")
    psi = load_chain(qb, str(tmp_path), L, "psiF")
    oc = qb.move_oc(psi, 0, L // 2 - 1)  # the reference leaves the centre at site 0 after whole sweeps
    assert oc == L // 2 - 1
    i = oc
    theta = psi[i].tensordot(psi[i + 1], [2], [0])
    assert max(theta.sizes()) >= 400
    # ---- truncated SVD of the real theta against the reference's svd(theta, 2, tol, min, max, pow) ----
    th_o = back(theta)
    orc.write_qtbt(th_o, str(tmp_path / "theta.qtbt"))
    f = lambda n: str(tmp_path / n)
    subprocess.run([HARNESS, "svdt", f("theta.qtbt"), "2", "1e-12", "4", str(D), "2.0", f("rU.qtbt"), f("rd.qtbt"), f("rV.qtbt"),
                    "--threads", THREADS], check=True, capture_output=True)
    rU, rd, rV = (orc.read_qtbt(f(n)) for n in ("rU.qtbt", "rd.qtbt", "rV.qtbt"))
    U, d, V = qb.svd(theta, 2, 1e-12, 4, D, 2.0)
    bU, bd, bV = back(U), back(d), back(V)
    assert orc.same_structure(bd, rd) and orc.same_structure(bU, rU) and orc.same_structure(bV, rV)
    assert orc.max_rel_err(bd, rd) <= 1e-12
    rec = back(U.mul_lastdim(d).tensordot(V.conj(), [2], [2]))
    assert orc.max_rel_err(rec, orc.recompose(rU, rd, rV)) <= 1e-11
    # ---- one Lanczos update of the same theta with the environments of that state ----
    mpo = [back(h) for h in H]
    mps = [back(p) for p in psi]
    lenv, renv = orc.trivial_edges(mps, mpo)
    El, Er = eng(qb, lenv), eng(qb, renv)
    for s in range(i):
        El = qb.compute_left_env(H[s], psi[s], El)
    for s in range(L - 1, i + 1, -1):
        Er = qb.compute_right_env(H[s], psi[s], Er)
    H2 = H[i].tensordot(H[i + 1], [2], [0]).permute([0, 1, 3, 4, 2, 5])
    for name, t in (("L", El), ("R", Er), ("H2", H2)):
        orc.write_qtbt(back(t), f(name + ".qtbt"))
    subprocess.run([HARNESS, "update", f("theta.qtbt"), f("H2.qtbt"), f("L.qtbt"), f("R.qtbt"), f("rE.qtbt"), f("rpsi.qtbt"),
                    "--threads", THREADS], check=True, capture_output=True)
    rE = orc.read_qtbt(f("rE.qtbt")).item()
    rpsi = orc.read_qtbt(f("rpsi.qtbt"))
    E, new = qb.two_sites_update(theta, H2, El, Er)
    assert abs(E - rE) <= 1e-10 * abs(rE)
    got = back(new)
    assert orc.same_structure(got, rpsi)
    assert orc.max_rel_err(got, rpsi) <= 1e-10


def test_dmrg_from_last_site_centre_and_logger_callback(engine):
    """reference dmrg_impl accepts an orthogonality centre on the last site (dmrg.cpp:229-233: stepped back by one without
    regauging) and reports every sweep to a dmrg_logger while the run is in progress (dmrg_logger.h)"""
    qb = engine
    from quantit_b200 import workloads as wl
    L = 10
    H = [qb.BTensor.from_host(**h) for h in wl.heisenberg_mpo(L)]
    mk = lambda: [qb.BTensor.from_host(**p_) for p_ in wl.random_mps(L, 4, L % 2, seed=3)]
    opts = qb.dmrg_options(1e-12, 1e-11, 32, 4, 40)
    psi0 = mk()
    E0 = qb.dmrg(H, psi0, opts, oc=0)
    psi1 = mk()
    oc = qb.move_oc(psi1, 0, L - 1)
    assert oc == L - 1
    seen = []
    E1, nsw, oc_out = qb.dmrg_logged(H, psi1, opts, lambda it, e, secs, bonds: seen.append((it, e, len(bonds))), oc=L - 1)
    assert oc_out == L - 2
    assert len(seen) == nsw and seen[-1][0] == nsw - 1 and seen[-1][2] == L + 1
    assert abs(seen[-1][1] - E1) <= 1e-12 * abs(E1)
    assert abs(E1 - E0) <= 1e-8 * abs(E0)  # the same ground state from either end
    assert abs(qb.contract(psi1, psi1, H) / qb.contract(psi1, psi1) - E1) <= 1e-8 * abs(E1)


def test_width6_cylinder_against_exact_diagonalisation(engine):
    """BASELINE.json configs[3] in miniature: Heisenberg cylinder of circumference 6 (2 x 6 = 12 sites, MPO bond up to 20
    sections before coalescing), U(1) two-site DMRG at a bond dimension that is exact for 12 sites, against the lowest
    eigenvalue of the dense Hamiltonian in the Sz = 0 sector"""
    qb = engine
    from quantit_b200 import workloads as wl
    Lx, Ly = 2, 6
    L = Lx * Ly
    H_sites = wl.heisenberg_cylinder_mpo(Lx, Ly)
    H = [qb.BTensor.from_host(**h) for h in H_sites]
    H = qb.coalesce(H, 1e-12)  # merge the equal-charge MPO states like the reference's bMPO::coalesce
    psi = [qb.BTensor.from_host(**p_) for p_ in wl.random_mps(L, 8, 0, seed=1)]
    E = qb.dmrg(H, psi, qb.dmrg_options(1e-14, 1e-12, 64, 4, 60), oc=0)
    # dense reference restricted to Sz = 0
    sz, up, I = np.diag([1.0, -1.0]), np.array([[0.0, 1.0], [0.0, 0.0]]), np.eye(2)
    dn = up.T

    def op(o, i):
        m = np.array([[1.0]])
        for k in range(L):
            m = np.kron(m, o if k == i else I)
        return m
    Hd = sum(0.25 * (2 * op(up, i) @ op(dn, j) + 2 * op(dn, i) @ op(up, j) + op(sz, i) @ op(sz, j))
             for i, j in wl.cylinder_bonds(Lx, Ly))
    mag = sum(np.diag(op(sz, i)) for i in range(L))
    keep = np.where(mag == 0)[0]
    E_exact = np.linalg.eigvalsh(Hd[np.ix_(keep, keep)])[0]
    assert abs(E - E_exact) <= 1e-9 * abs(E_exact), (E, E_exact)
