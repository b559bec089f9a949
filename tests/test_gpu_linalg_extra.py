"""GPU parity tests of the round-2 entry points: block eigh (+ corrected truncation), the free-standing truncate,
reshape / reshape_as, tensorgdot (fused epilogue and union fallback), packed tensor files.

eigh has no reference behaviour to be in parity with (the reference's implementation crashes in the oracle build, see
oracle/ref_harness.cpp `eigh`): the contract is the mathematical one, checked against numpy.linalg.eigh per charge group
(eigenvalues 1e-12 of the largest, ascending inside a group) and through A = U diag(e) U^T, U^T U = 1."""
import numpy as np
import pytest

import qtb_oracle as orc
from quantit_b200 import workloads as wl

pytestmark = pytest.mark.gpu


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


def symmetric_matrix(rng, sizes, charges):
    """rank-2 block matrix on (leg, conj leg), selection rule 0, symmetric"""
    leg = ([list(sizes)], [[(q,) for q in charges]])
    t = orc.BT([list(sizes), list(sizes)], [[(q,) for q in charges], [(-q,) for q in charges]], (0,), {})
    for i in range(len(sizes)):
        for j in range(i, len(sizes)):
            if charges[i] == charges[j] and rng.random() < 0.8 or i == j:
                if charges[i] != charges[j]:
                    continue
                blk = rng.standard_normal((sizes[i], sizes[j]))
                if i == j:
                    blk = blk + blk.T
                t.blocks[(i, j)] = blk
                t.blocks[(j, i)] = blk.T.copy()
    return t


@pytest.mark.parametrize("sizes,charges", [([3, 2, 4], [0, 1, 2]), ([3, 2, 4, 5, 1], [0, 1, 0, 2, 1]),
                                           ([40, 70, 55, 30], [0, 0, 1, 1]), ([500, 420], [0, 0])])
def test_eigh_matches_numpy_per_group(engine, sizes, charges):
    rng = np.random.default_rng(5)
    m = symmetric_matrix(rng, sizes, charges)
    e, U = engine.eigh(eng(engine, m), 1)
    be, bU = back(e), back(U)
    groups = orc.eigh_groups(m, 1)
    assert be.sec_sizes == [[len(g[0]) for g in groups]]
    scale = max(float(np.max(np.abs(g[0]))) for g in groups)
    for b_i, (ev, _, rows, _) in enumerate(groups):
        got = be.blocks[(b_i,)]
        assert np.all(np.diff(got) >= -1e-12 * scale), "eigenvalues must ascend inside a group"
        assert np.max(np.abs(got - ev)) <= 1e-12 * scale
        # A_g = U diag(e) U^T and U^T U = 1, from the blocks of U of this group
        Ug = np.concatenate([bU.blocks[(sec, b_i)] for sec, _ in rows], axis=0)
        dense = orc.svd_groups(m)[b_i][0]
        assert np.max(np.abs(Ug @ np.diag(got) @ Ug.T - dense)) <= 1e-11 * scale
        assert np.max(np.abs(Ug.T @ Ug - np.eye(Ug.shape[1]))) <= 1e-11
    # the same through the engine's own ops: (U * e) . U^T over the bond reproduces the matrix
    rec = back(U.mul_lastdim(e).tensordot(U.conj(), [1], [1]))
    for k, v in m.blocks.items():
        assert np.max(np.abs(rec.blocks[k] - v)) <= 1e-11 * scale


def test_eigh_two_site_density_matrix_and_truncation(engine):
    """rho = theta theta^T over the right legs: rank 4, block symmetric, positive semi-definite; eigh(rho, 2) and the
    corrected truncation (threshold over |e|)"""
    rng = np.random.default_rng(9)
    beta = wl.bond(5, 40, 1.2, 2)
    th = wl.rand_like(wl.shape([beta, wl.SPIN_HALF, wl.SPIN_HALF, wl.conj_leg(beta)], (0,)), rng)
    theta = orc.BT(**{k: th[k] for k in ("sec_sizes", "cvals", "sel", "blocks")})
    rho = orc.tensordot(theta, orc.conj(theta), [2, 3], [2, 3])
    R = eng(engine, rho)
    e, U = engine.eigh(R, 2)
    groups = orc.eigh_groups(rho, 2)
    be = back(e)
    scale = max(float(np.max(np.abs(g[0]))) for g in groups)
    for b_i, g in enumerate(groups):
        assert np.max(np.abs(be.blocks[(b_i,)] - g[0])) <= 1e-12 * scale
    # eigenvalues of rho are the squared singular values of theta
    _, d, _ = engine.svd(eng(engine, theta), 2)
    sv2 = np.sort(np.concatenate([v.ravel() for v in back(d).blocks.values()]) ** 2)
    ev = np.sort(np.concatenate([v.ravel() for v in be.blocks.values()]))
    assert np.max(np.abs(ev[-len(sv2):] - sv2)) <= 1e-12 * scale
    # truncated: keep at most 12 eigenpairs, the ones of largest |e|
    et, Ut = engine.eigh(R, 2, 0.0, 1, 12, 1.0)
    kept = np.sort(np.concatenate([v.ravel() for v in back(et).blocks.values()]))
    assert len(kept) == 12 and np.max(np.abs(kept - ev[-12:])) <= 1e-12 * scale
    UtU = back(Ut.conj().tensordot(Ut, [0, 1], [0, 1]))
    for (i, j), blk in UtU.blocks.items():
        assert i == j and np.allclose(blk, np.eye(blk.shape[0]), atol=1e-11)


def test_eigh_rejects_non_square_groups(engine):
    t = orc.BT([[2, 3], [4]], [[(0,), (1,)], [(0,)]], (0,), {(0, 0): np.ones((2, 4))})
    with pytest.raises(engine.InvalidArgument):
        engine.eigh(eng(engine, t), 1)


@pytest.mark.parametrize("args", [(1e-1, 1, 1000, 2.0), (1e-4, 4, 18, 2.0), (0.0, 1, 7, 2.0), (3.0, 1, 1000, 2.0)])
def test_truncate_free_function_matches_truncated_svd(engine, args):
    """truncate(U, d, V, max, min, tol, pow) applied to an untruncated svd == svd(A, split, tol, min, max, pow)"""
    rng = np.random.default_rng(21)
    theta = orc.BT(**{k: v for k, v in wl.random_btensor(rng, 4, max_sec=4, max_size=6).items()
                      if k in ("sec_sizes", "cvals", "sel", "blocks")})
    T = eng(engine, theta)
    tol, mn, mx, pw = args
    U, d, V = engine.svd(T, 2)
    Ut, dt, Vt = engine.truncate(U, d, V, mx, mn, tol, pw)
    oU, od, oV = orc.svd_trunc(theta, 2, tol, mn, mx, pw)
    for got, want in ((dt, od), (Ut, oU), (Vt, oV)):
        assert orc.same_structure(back(got), want)
    assert orc.max_rel_err(back(dt), od) <= 1e-12
    rec = Ut.mul_lastdim(dt).tensordot(Vt.conj(), [Ut.dim() - 1], [Vt.dim() - 1])
    assert_same(back(rec), orc.recompose(oU, od, oV), 1e-11)


@pytest.mark.parametrize("seed", range(6))
def test_reshape_and_reshape_as(engine, seed):
    rng = np.random.default_rng(300 + seed)
    rank = int(rng.integers(2, 6))
    t = orc.BT(**{k: v for k, v in wl.random_btensor(rng, rank).items() if k in ("sec_sizes", "cvals", "sel", "blocks")})
    T = eng(engine, t)
    cuts = sorted(int(x) for x in rng.choice(np.arange(0, rank + 1), size=int(rng.integers(0, 3)), replace=True))
    R = T.reshape(cuts)
    assert_same(back(R), orc.reshape(t, cuts), 0.0)
    # on a permuted (strided) view as well: the blocks are gathered first
    perm = [int(x) for x in rng.permutation(rank)]
    assert_same(back(T.permute(perm).reshape(cuts)), orc.reshape(orc.permute(t, perm), cuts), 0.0)
    # reshape_as brings it back (both modes)
    assert_same(back(R.reshape_as(T)), t, 0.0)
    assert_same(back(R.reshape_as(T, overwrite_c_vals=True)), t, 0.0)


def test_reshape_as_rejects_incompatible(engine):
    a = orc.BT([[2, 3], [4, 1]], [[(0,), (1,)], [(0,), (-1,)]], (0,), {(0, 0): np.ones((2, 4)), (1, 1): np.ones((3, 1))})
    b = orc.BT([[2, 3, 1], [4, 1]], [[(0,), (1,), (0,)], [(0,), (-1,)]], (0,), {})
    with pytest.raises(engine.InvalidArgument):
        eng(engine, a).reshape_as(eng(engine, b))


@pytest.mark.parametrize("n_sec,D,sigma", [(5, 24, 1.0), (9, 96, 1.5), (13, 300, 2.0)])
def test_tensorgdot_epilogue_and_union(engine, n_sec, D, sigma):
    a, b, da, db = wl.tdot_pair(n_sec, D, sigma, seed=3)
    oa, ob = orc.BT(**a), orc.BT(**b)
    oc = orc.tensordot(oa, ob, da, db)
    rng = np.random.default_rng(1)
    c_full = oc.structure_like()
    for k, v in oc.blocks.items():
        c_full.blocks[k] = rng.standard_normal(v.shape)
    A, B = eng(engine, oa), eng(engine, ob)
    launches0 = engine.default_context().counters()["kernel_launches"]
    got = eng(engine, c_full).tensorgdot(A, B, da, db, beta=0.75, alpha=-1.5)
    # same block table as A.B: ONE kernel (the grouped GEMM with the linear combination in its epilogue)
    assert engine.default_context().counters()["kernel_launches"] - launches0 == 1
    assert_same(back(got), orc.tensorgdot(c_full, oa, ob, da, db, 0.75, -1.5), 1e-12)
    # C with a different block list (every other block missing): union of the two lists
    c_part = oc.structure_like()
    for n, (k, v) in enumerate(sorted(oc.blocks.items())):
        if n % 2 == 0:
            c_part.blocks[k] = rng.standard_normal(v.shape)
    got = eng(engine, c_part).tensorgdot(A, B, da, db, beta=2.0, alpha=0.5)
    assert_same(back(got), orc.tensorgdot(c_part, oa, ob, da, db, 2.0, 0.5), 1e-12)


def test_tensor_file_round_trip(engine, tmp_path):
    rng = np.random.default_rng(77)
    t = orc.BT(**{k: v for k, v in wl.random_btensor(rng, 3, nc=2).items() if k in ("sec_sizes", "cvals", "sel", "blocks")})
    T = eng(engine, t)
    path = str(tmp_path / "t.qtbpack")
    T.permute([2, 0, 1]).save(path)
    assert_same(back(engine.BTensor.load(path)), orc.permute(t, [2, 0, 1]), 0.0)
    with pytest.raises(engine.QtbError):
        engine.BTensor.load(str(tmp_path / "missing.qtbpack"))
