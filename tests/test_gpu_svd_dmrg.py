"""GPU parity tests: block SVD + truncation, Lanczos update and environment updates vs the oracle and the committed
reference outputs (tests/golden). Block structure must be identical; singular values 1e-10 relative to the largest
(north star), in practice ~1e-14; singular vectors are gauge dependent and are compared through U d V^T."""
import os

import numpy as np
import pytest

import qtb_oracle as orc
from quantit_b200 import workloads as wl

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def g(name):
    return orc.read_qtbt(os.path.join(G, name + ".qtbt"))


def eng(qb, t: orc.BT):
    return qb.BTensor.from_host(t.sec_sizes, t.cvals, t.sel, t.blocks, mods=t.mods)


def back(t) -> orc.BT:
    ss, cv, sel, mods = t.structure()
    return orc.BT(ss, cv, sel, t.to_host(), None if not any(mods) else mods)


def assert_same(got: orc.BT, want: orc.BT, tol):
    assert got.sec_sizes == want.sec_sizes and got.cvals == want.cvals and got.sel == want.sel
    assert sorted(got.blocks) == sorted(want.blocks)
    for k in want.blocks:
        assert got.blocks[k].shape == want.blocks[k].shape, k
    assert orc.max_rel_err(got, want) <= tol


def check_svd(qb, theta: orc.BT, split, args, ref=None):
    T = eng(qb, theta)
    if args is None:
        U, d, V = qb.svd(T, split)
        oU, od, oV = ref or orc.svd(theta, split)
    else:
        U, d, V = qb.svd(T, split, *args)
        oU, od, oV = ref or orc.svd_trunc(theta, split, *args)
    bU, bd, bV = back(U), back(d), back(V)
    assert orc.same_structure(bd, od), (bd.sec_sizes, od.sec_sizes, sorted(bd.blocks), sorted(od.blocks))
    assert orc.same_structure(bU, oU) and orc.same_structure(bV, oV)
    assert orc.max_rel_err(bd, od) <= 1e-12
    # gauge-independent comparison through the engine's own ops: (U*d) . V^T over the bond
    rec = U.mul_lastdim(d).tensordot(V.conj(), [U.dim() - 1], [V.dim() - 1])
    assert_same(back(rec), orc.recompose(oU, od, oV), 1e-11)
    # orthonormal columns: U^T U = 1 on every kept bond sector
    UtU = back(U.conj().tensordot(U, list(range(U.dim() - 1)), list(range(U.dim() - 1))))
    for (i, j), blk in UtU.blocks.items():
        assert i == j
        nz = np.linalg.norm(blk, axis=0) > 0.5  # exact-zero singular directions have no vector (documented)
        assert np.allclose(blk[np.ix_(nz, nz)], np.eye(int(nz.sum())), atol=1e-11)


@pytest.mark.parametrize("tag,args", [("svd", None), ("svdt", (1e-4, 4, 18, 2.0)), ("svdt2", (1e-1, 1, 1000, 2.0)),
                                       ("svdt3", (0.5, 1, 1000, 2.0))])
def test_svd_golden(engine, tag, args):
    check_svd(engine, g("svd_theta"), 2, args, ref=(g(tag + "_U"), g(tag + "_d"), g(tag + "_V")))


@pytest.mark.parametrize("seed,tol,mn,mx", [(0, 1e-3, 1, 10 ** 6), (1, 0.3, 1, 10 ** 6), (2, 0.0, 1, 7), (3, 0.8, 2, 10 ** 6),
                                            (4, 1e-12, 4, 12), (5, 1e-8, 4, 40)])
def test_svd_trunc_random(engine, seed, tol, mn, mx):
    th = wl.rand_like(wl.shape([wl.bond(7, 30 + seed, 1.5, 2), wl.SPIN_HALF, wl.SPIN_HALF,
                                wl.conj_leg(wl.bond(7, 26, 1.5, 2))], (0,)), np.random.default_rng(seed))
    for k in th["blocks"]:
        blk = th["blocks"][k]
        u, s, vt = np.linalg.svd(blk.reshape(blk.shape[0], -1), full_matrices=False)
        th["blocks"][k] = ((u * (s * np.exp(-1.0 * np.arange(len(s))))) @ vt).reshape(blk.shape)
    check_svd(engine, orc.BT(**th), 2, (tol, mn, mx, 2.0))


def test_svd_larger_groups_and_splits(engine):
    """groups wider than one Jacobi column block, tall and wide groups, split 1 and 3, ZxZ charges with missing blocks"""
    rng = np.random.default_rng(77)
    th = wl.rand_like(wl.shape([wl.bond(5, 150, 1.2, 2), wl.SPIN_HALF, wl.SPIN_HALF, wl.conj_leg(wl.bond(5, 90, 1.2, 2))],
                               (0,)), rng)
    for split in (1, 2, 3):
        check_svd(engine, orc.BT(**th), split, None)
    t2 = wl.random_btensor(rng, 4, max_sec=3, max_size=6, nc=2, fill=0.7)
    check_svd(engine, orc.BT(**t2), 2, None)
    check_svd(engine, orc.BT(**t2), 2, (1e-2, 1, 5, 2.0))


@pytest.mark.parametrize("graded", [False, True])
@pytest.mark.parametrize("bonds", [(700, 620), (1500, 1300)])
def test_svd_large_panel_path(engine, graded, bonds):
    """bonds (700, 620): groups of ~450 x 400, (m + n) <= 879 rows -> shared-memory panels of 32 columns (block Jacobi
    with 16-wide column blocks). bonds (1500, 1300): groups of ~950 x 830 -> panels that do not fit in shared memory ->
    the tensor-core path (DMMA Gram / two-sided Jacobi eig with descending eigenvalue order / DMMA update, the charge
    groups dealt to lanes on separate CUDA streams). Random blocks and DMRG-like graded blocks (singular values spanning
    40 orders of magnitude and more, numerically rank deficient)."""
    rng = np.random.default_rng(5)
    th = wl.rand_like(wl.shape([wl.bond(5, bonds[0], 1.2, 2), wl.SPIN_HALF, wl.SPIN_HALF,
                                wl.conj_leg(wl.bond(5, bonds[1], 1.2, 2))], (0,)), rng)
    if graded:
        for k in th["blocks"]:
            blk = th["blocks"][k]
            u, s, vt = np.linalg.svd(blk.reshape(blk.shape[0], -1), full_matrices=False)
            th["blocks"][k] = ((u * (s * np.exp(-0.4 * np.arange(len(s))))) @ vt).reshape(blk.shape)
        check_svd(engine, orc.BT(**th), 2, (1e-10, 4, 300, 2.0))
    else:
        check_svd(engine, orc.BT(**th), 2, None)
        check_svd(engine, orc.BT(**th), 2, (1e-3, 4, 500, 2.0))


def test_heff_env_update_golden(engine):
    qb = engine
    psi, H2, L, R, W = (eng(qb, g("heff_" + n)) for n in ("psi", "H2", "L", "R", "W"))
    assert_same(back(qb.hamil2site_times_state(psi, H2, L, R)), g("heff_out"), 1e-13)
    assert_same(back(qb.compute_left_env(W, eng(qb, g("env_Y")), L)), g("env_left"), 1e-13)
    assert_same(back(qb.compute_right_env(W, eng(qb, g("env_Y2")), R)), g("env_right"), 1e-13)
    E, p = qb.two_sites_update(psi, H2, L, R)
    assert abs(E - g("update_E").item()) <= 1e-12 * abs(E)
    assert_same(back(p), g("update_psi"), 1e-12)


@pytest.mark.parametrize("name", ["dmrg_heis8", "dmrg_hub4"])
def test_dmrg_against_reference_run(engine, name):
    """whole two-site DMRG (U(1) Heisenberg L=8; U(1)xU(1) Hubbard L=4): same MPO and initial MPS as the reference run
    (dumped by oracle/_ref/ref_harness), energies per sweep to 1e-10 relative, identical final block structure"""
    import json
    qb = engine
    d = os.path.join(G, name)
    rec = json.load(open(os.path.join(d, "reference_run.json")))
    L = rec["L"]
    H = [eng(qb, orc.read_qtbt(os.path.join(d, f"H_{i}.qtbt"))) for i in range(L)]
    psi = [eng(qb, orc.read_qtbt(os.path.join(d, f"psi0_{i}.qtbt"))) for i in range(L)]
    log = {}
    opt = qb.dmrg_options(rec["cutoff"], rec["convergence_criterion"], rec["maximum_bond"], rec["minimum_bond"],
                          rec["maximum_iterations"])
    E = qb.dmrg(H, psi, opt, oc=rec["oc"], log=log)
    assert len(log["energy"]) == len(rec["sweep_energy"])
    assert np.allclose(log["energy"], rec["sweep_energy"], rtol=1e-10, atol=0), (log["energy"], rec["sweep_energy"])
    assert abs(E - rec["E0"]) <= 1e-10 * abs(rec["E0"])
    assert log["mid_bond"] == rec["mid_bond"]
    for i in range(L):
        want = orc.read_qtbt(os.path.join(d, f"psiF_{i}.qtbt"))
        got = back(psi[i])
        assert got.sec_sizes == want.sec_sizes and got.cvals == want.cvals and got.sel == want.sel, i
        assert sorted(got.blocks) == sorted(want.blocks), i


# ---- bMPS chain contractions and gauge moves (reference sources/MPT.cpp:75-111, 211-233, 275-292) ------------------------
def _chain(name, prefix):
    import json
    d = os.path.join(G, name)
    rec = json.load(open(os.path.join(d, "reference_run.json")))
    L = rec["L"]
    return rec, [orc.read_qtbt(os.path.join(d, f"H_{i}.qtbt")) for i in range(L)], \
        [orc.read_qtbt(os.path.join(d, f"{prefix}_{i}.qtbt")) for i in range(L)]


@pytest.mark.parametrize("name", ["dmrg_heis8", "dmrg_hub4"])
def test_contract_and_move_oc_against_reference_run(engine, name):
    """contract(psi, psi[, H]) of the reference's final MPS against the values the reference printed; bMPS::move_oc
    against the reference's own moved chains (bit-exact structure per site, overlap with the reference's chain = 1 to 1e-12)"""
    qb = engine
    rec, Ho, psio = _chain(name, "psiF")
    L = rec["L"]
    H = [eng(qb, t) for t in Ho]
    psi = [eng(qb, t) for t in psio]
    assert abs(qb.contract(psi, psi, H) - rec["contract_E"]) <= 1e-12 * abs(rec["contract_E"])
    assert abs(qb.contract(psi, psi) - rec.get("norm", 1.0)) <= 1e-12

    def check(prefix):
        want = _chain(name, prefix)[2]
        got = [back(t) for t in psi]
        for a, b in zip(got, want):
            assert a.sec_sizes == b.sec_sizes and a.cvals == b.cvals and a.sel == b.sel
            assert sorted(a.blocks) == sorted(b.blocks)
        # the site tensors are defined up to a sign per bond index; the chains must be the same normalised state
        assert abs(qb.contract(psi, [eng(qb, t) for t in want]) - 1.0) <= 1e-12
        assert abs(orc.contract(got, want) - 1.0) <= 1e-12

    oc = qb.move_oc(psi, rec["oc"], L - 2)
    assert oc == L - 2
    check("psiM")
    oc = qb.move_oc(psi, oc, 1)
    assert oc == 1
    check("psiN")
    assert abs(qb.contract(psi, psi, H) - rec["contract_E"]) <= 1e-11 * abs(rec["contract_E"])
    assert abs(qb.contract(psi, psi) - 1.0) <= 1e-12
    with pytest.raises(qb.InvalidArgument):
        qb.move_oc(psi, oc, L)


@pytest.mark.parametrize("name", ["dmrg_heis8", "dmrg_hub4"])
def test_coalesce_against_reference_run(engine, name):
    """bMPO::coalesce(1e-10) against the reference's own coalesced MPO: bit-exact structure per site, the operator through
    gauge-invariant matrix elements <a|H|b> on the run's initial and final states (1e-12)"""
    qb = engine
    rec, Ho, psiFo = _chain(name, "psiF")
    L = rec["L"]
    Hco = [orc.read_qtbt(os.path.join(G, name, f"Hc_{i}.qtbt")) for i in range(L)]
    psi0o = _chain(name, "psi0")[2]
    H = qb.coalesce([eng(qb, t) for t in Ho], 1e-10)
    for a, b in zip([back(t) for t in H], Hco):
        assert a.sec_sizes == b.sec_sizes and a.cvals == b.cvals and a.sel == b.sel
        assert sorted(a.blocks) == sorted(b.blocks)
    psiF, psi0 = [eng(qb, t) for t in psiFo], [eng(qb, t) for t in psi0o]
    for (x, y), (xo, yo) in zip(((psiF, psiF), (psi0, psiF), (psi0, psi0)), ((psiFo, psiFo), (psi0o, psiFo), (psi0o, psi0o))):
        want = orc.contract(xo, yo, Hco)
        assert abs(qb.contract(x, y, H) - want) <= 1e-12 * max(1.0, abs(want))
