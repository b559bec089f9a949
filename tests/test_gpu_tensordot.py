"""GPU parity tests of the contraction path: engine (C ABI -> CUDA) vs the oracle on the same seeded inputs.

Mirrors the reference's own tensordot tests (include/blockTensor/btensor.h:1551-1628): values vs a per-block
reference, index-order invariance of a full trace, outer product through empty dims; plus the block-structure
bit-exactness the north star asks for. Tolerance: 1e-12 norm-wise relative (fp64; north star allows 1e-10).
"""
import itertools

import numpy as np
import pytest

import qtb_oracle as orc
from quantit_b200 import workloads as wl

pytestmark = pytest.mark.gpu
TOL = 1e-12


def to_engine(qb, d):
    return qb.BTensor.from_host(**{k: d[k] for k in ("sec_sizes", "cvals", "sel", "blocks")})


def to_oracle(d):
    return orc.BT(**{k: d[k] for k in ("sec_sizes", "cvals", "sel", "blocks")})


def check_same(got_t, want: orc.BT, tol=TOL):
    ss, cv, sel, _ = got_t.structure()
    assert ss == want.sec_sizes
    assert cv == want.cvals
    assert sel == want.sel
    got = got_t.to_host()
    assert sorted(got) == sorted(want.blocks), "block list differs"
    assert list(got) == sorted(want.blocks), "block order is not lexicographic"
    den = max([float(np.max(np.abs(v))) for v in want.blocks.values() if v.size] + [0.0])
    for k, v in want.blocks.items():
        assert got[k].shape == v.shape
        if v.size:
            assert float(np.max(np.abs(got[k] - v))) <= tol * max(den, 1e-300), k


@pytest.mark.parametrize("n_sec,D,sigma", [(5, 24, 1.0), (9, 96, 1.5), (13, 300, 2.0), (41, 2048, 6.0)])
def test_tdot_rank4_one_bond(engine, n_sec, D, sigma):
    a, b, da, db = wl.tdot_pair(n_sec, D, sigma, seed=11)
    C = to_engine(engine, a).tensordot(to_engine(engine, b), da, db)
    check_same(C, orc.tensordot(to_oracle(a), to_oracle(b), da, db))


def test_tdot_flops_match_oracle(engine):
    a, b, da, db = wl.tdot_pair(41, 2048, 6.0, seed=1234)
    info = to_engine(engine, a).tensordot_info(to_engine(engine, b), da, db)
    assert info["flops"] == orc.tensordot_flops(to_oracle(a), to_oracle(b), da, db)
    assert info["out_blocks"] == 642 and info["pairs"] == 642


@pytest.mark.parametrize("seed", range(12))
def test_tdot_random_structures(engine, seed):
    """random ranks / sections / charges / missing blocks, multi-index contractions in arbitrary order"""
    rng = np.random.default_rng(100 + seed)
    ra, rb = int(rng.integers(1, 5)), int(rng.integers(1, 5))
    k = int(rng.integers(0, min(ra, rb) + 1))
    a = wl.random_btensor(rng, ra, nc=1 + seed % 2)
    dims_a = [int(x) for x in rng.permutation(ra)[:k]]
    dims_b = [int(x) for x in rng.permutation(rb)[:k]]
    legs_b = [None] * rb
    for da_, db_ in zip(dims_a, dims_b):
        legs_b[db_] = wl.conj_leg((a["sec_sizes"][da_], a["cvals"][da_]))
    nc = len(a["sel"])
    for i in range(rb):
        if legs_b[i] is None:
            ns = int(rng.integers(1, 4))
            legs_b[i] = ([int(x) for x in rng.integers(1, 5, ns)],
                         [tuple(int(x) for x in rng.integers(-2, 3, nc)) for _ in range(ns)])
    b = wl.random_btensor(rng, rb, nc=nc, legs=legs_b)
    C = to_engine(engine, a).tensordot(to_engine(engine, b), dims_a, dims_b)
    check_same(C, orc.tensordot(to_oracle(a), to_oracle(b), dims_a, dims_b))


def test_tdot_on_permuted_views(engine):
    """permute() is metadata-only in the engine; the strided gather is fused into the GEMM operand load"""
    rng = np.random.default_rng(5)
    a = wl.random_btensor(rng, 4, max_sec=3, max_size=6, fill=0.9)
    A, oA = to_engine(engine, a), to_oracle(a)
    for perm in [(3, 1, 0, 2), (1, 0, 3, 2), (2, 3, 0, 1)]:
        Ap, oAp = A.permute(perm), orc.permute(oA, perm)
        check_same(Ap, oAp)
        Bc, oBc = Ap.conj(), orc.conj(oAp)
        check_same(Bc, oBc)
        for dims in [([0], [0]), ([1, 3], [1, 3]), ([0, 1, 2, 3], [0, 1, 2, 3]), ([2, 0], [2, 0])]:
            check_same(Ap.tensordot(Bc, *dims), orc.tensordot(oAp, oBc, *dims))


def test_trace_invariant_under_index_order(engine):
    """reference btensor.h:1551-1580: full contraction of A with conj(A) under every index order gives the same rank-0
    tensor with the single block index ()"""
    rng = np.random.default_rng(9)
    a = wl.random_btensor(rng, 3, max_sec=3, max_size=5, fill=1.0)
    A = to_engine(engine, a)
    Ac = A.conj()
    want = sum(float(np.sum(v * v)) for v in a["blocks"].values())
    for perm in itertools.permutations(range(3)):
        r = A.tensordot(Ac, list(perm), list(perm))
        assert r.dim() == 0 and r.block_indices() == [()]
        assert abs(r.item() - want) <= 1e-12 * abs(want)


def test_outer_product_empty_dims(engine):
    """reference btensor.h:1617-1628: no contracted dims -> every block pair yields a block"""
    rng = np.random.default_rng(10)
    a = wl.random_btensor(rng, 2, max_sec=3, fill=1.0)
    b = wl.random_btensor(rng, 2, max_sec=3, fill=1.0)
    C = to_engine(engine, a).tensordot(to_engine(engine, b), [], [])
    want = orc.tensordot(to_oracle(a), to_oracle(b), [], [])
    assert len(want.blocks) == len(a["blocks"]) * len(b["blocks"])
    check_same(C, want)


def test_tdot_errors(engine):
    """error behaviour of compute_tdot_shape / check_product_compat (btensor.cpp:783-883): TORCH_CHECK -> CheckError"""
    rng = np.random.default_rng(3)
    a = wl.random_btensor(rng, 2)
    A = to_engine(engine, a)
    with pytest.raises(engine.CheckError):
        A.tensordot(A, [0], [0, 1])
    with pytest.raises(engine.CheckError):  # charges not pairwise inverse (unless all neutral)
        bad = dict(a)
        bad["cvals"] = [[(q[0] + 1,) for q in cs] for cs in a["cvals"]]
        bad["sel"] = (a["sel"][0] + 2,)
        to_engine(engine, bad).tensordot(A, [0], [0])
    with pytest.raises(engine.InvalidArgument):  # block violating the selection rule (check_tensor)
        bad = dict(a)
        bad["sel"] = (a["sel"][0] + 7,)
        if not bad["blocks"]:
            raise engine.InvalidArgument("no blocks")
        to_engine(engine, bad)


@pytest.mark.parametrize("mpo_sizes", [[1, 1, 1, 1, 1], [3, 2, 1, 2, 3], [5, 7, 4]])
def test_tdot_with_small_mpo_legs(engine, mpo_sizes):
    """the MPO step of H_eff.psi / the environment updates: a big intermediate [w, a', s, b] contracted with a tensor
    whose dims are all small — the skinny streaming kernel (K, N <= 16) — with MPO bond sections of size 1 (Heisenberg,
    Hubbard) and of size > 1 (a coalesced MPO: K and N > 1, row offsets no longer a single stride)."""
    rng = np.random.default_rng(31)
    n = len(mpo_sizes)
    charges = [(0,), (-2,), (2,), (0,), (0,)][:n]
    omega = (mpo_sizes, charges)
    beta = wl.bond(7, 180, 1.4, 2)
    beta_odd = (beta[0], [(c[0] - 1,) for c in beta[1]])  # the bond on the other side of one spin-1/2 site
    t1 = wl.rand_like(wl.shape([omega, wl.conj_leg(beta), wl.SPIN_HALF, beta_odd], (0,)), rng)
    W = wl.rand_like(wl.shape([wl.conj_leg(omega), wl.SPIN_HALF, omega, wl.conj_leg(wl.SPIN_HALF)], (0,)), rng)
    A, B = to_engine(engine, t1), to_engine(engine, W)
    C = A.tensordot(B, [0, 2], [0, 3])
    check_same(C, orc.tensordot(to_oracle(t1), to_oracle(W), [0, 2], [0, 3]))
    info = A.tensordot_info(B, [0, 2], [0, 3])
    assert info["flops"] == orc.tensordot_flops(to_oracle(t1), to_oracle(W), [0, 2], [0, 3])


def test_heff_and_envs(engine):
    """H_eff.psi and the environment updates are chains of 3 fused contractions (dmrg.cpp:424-531)"""
    qb = engine
    psi, W, L, R = wl.heff_set(7, 48, 1.3, seed=3)
    Wb = to_engine(qb, W)
    H2 = Wb.tensordot(Wb, [2], [0]).permute([0, 1, 3, 4, 2, 5])
    oW = to_oracle(W)
    oH2 = orc.permute(orc.tensordot(oW, oW, [2], [0]), [0, 1, 3, 4, 2, 5])
    check_same(H2, oH2)
    out = qb.hamil2site_times_state(to_engine(qb, psi), H2, to_engine(qb, L), to_engine(qb, R))
    t = orc.tensordot(to_oracle(L), to_oracle(psi), [0], [0])
    t = orc.tensordot(t, oH2, [0, 2, 3], [0, 4, 5])
    t = orc.tensordot(t, to_oracle(R), [1, 4], [0, 1])
    check_same(out, t)


@pytest.mark.parametrize("seed", range(6))
def test_add_matches_reference_merge(engine, seed):
    """btensor::add(other, alpha) including the reference's flat_map::merge behaviour (see include/qtb.h qtb_add)"""
    rng = np.random.default_rng(700 + seed)
    sh = wl.random_btensor(rng, int(rng.integers(1, 4)), max_sec=4, fill=1.0)
    keys = list(sh["blocks"])
    a = dict(sh, blocks={k: sh["blocks"][k] for k in keys if rng.random() < 0.6})
    b = dict(sh, blocks={k: rng.standard_normal(sh["blocks"][k].shape) for k in keys if rng.random() < 0.6})
    alpha = float(rng.choice([-1.0, 0.5, 2.0]))
    check_same(to_engine(engine, a).add(to_engine(engine, b), alpha), orc.add(to_oracle(a), to_oracle(b), alpha))
