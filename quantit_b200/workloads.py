"""Synthetic block-sparse workloads of BASELINE.json's configs, as plain host descriptions.

A host description is a dict {sec_sizes, cvals, sel, blocks}: section sizes per dim, section charges per dim (tuples of
ints), selection rule, and {block index: C-contiguous fp64 ndarray}. It converts to the engine's BTensor with
`BTensor.from_host(**desc)` and (in tests) to the oracle's BT with `BT(**desc)`.

Values are uniform [0,1) from a seeded numpy Generator (the reference's rand_like fills blocks from the global torch
RNG, sources/btensor.cpp:2399-2404; inputs are exchanged as explicit arrays instead of re-deriving that stream).
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Sequence, Tuple

import numpy as np

Charge = Tuple[int, ...]


def bond(n_sec: int, D: int, sigma: float, step: int = 2):
    """SURVEY.md §8d bond generator: sectors i with charge step*(i - n_sec//2), Gaussian weights, sizes
    max(1, round(D*w_i/sum w)), remainder added to the centre sector so that the sizes sum to D."""
    c = n_sec // 2
    w = np.exp(-((np.arange(n_sec) - c) ** 2) / (2.0 * sigma * sigma))
    sizes = np.maximum(1, np.rint(D * w / w.sum()).astype(np.int64))
    sizes[c] += D - int(sizes.sum())
    assert sizes[c] >= 1 and int(sizes.sum()) == D
    charges = [(int(step * (i - c)),) for i in range(n_sec)]
    return [int(s) for s in sizes], charges


def conj_leg(leg):
    sizes, charges = leg
    return list(sizes), [tuple(-x for x in q) for q in charges]


SPIN_HALF = ([1, 1], [(1,), (-1,)])                      # sigma of SURVEY.md §8d
HEIS_MPO_BOND = ([1, 1, 1, 1, 1], [(0,), (-2,), (2,), (0,), (0,)])   # omega of §8d / appendix C


def shape(legs, sel: Charge):
    return {"sec_sizes": [list(l[0]) for l in legs], "cvals": [list(l[1]) for l in legs], "sel": tuple(sel),
            "blocks": {}}


def allowed_indices(desc) -> List[Tuple[int, ...]]:
    nc = len(desc["sel"])
    out = []
    ranges = [range(len(s)) for s in desc["sec_sizes"]]
    for idx in itertools.product(*ranges):
        q = tuple(sum(desc["cvals"][d][i][c] for d, i in enumerate(idx)) for c in range(nc))
        if q == tuple(desc["sel"]):
            out.append(idx)
    return out


def rand_like(desc, rng: np.random.Generator):
    out = dict(desc)
    out["blocks"] = {}
    for idx in allowed_indices(desc):
        dims = tuple(desc["sec_sizes"][d][i] for d, i in enumerate(idx))
        out["blocks"][idx] = rng.random(dims)
    return out


def tdot_pair(n_sec: int, D: int, sigma: float, seed: int = 1234):
    """T1 / T2 of SURVEY.md §8d: A, B = rand_like(shape_from(beta, s, s, beta.conj())); C = A.tensordot(B,{3},{0})."""
    rng = np.random.default_rng(seed)
    beta = bond(n_sec, D, sigma, 2)
    sh = shape([beta, SPIN_HALF, SPIN_HALF, conj_leg(beta)], (0,))
    return rand_like(sh, rng), rand_like(sh, rng), [3], [0]


T1 = dict(n_sec=41, D=2048, sigma=6.0)    # BASELINE.json configs[1] ("~40 sectors, total dim 2048")
T2 = dict(n_sec=15, D=4096, sigma=1.6)    # DMRG-like sector profile at D=4096
T2_8K = dict(n_sec=17, D=8192, sigma=1.8)


def heisenberg_W(Jp: float = 0.25):
    """Bulk U(1) Heisenberg MPO site W[wl, s', wr, s] of the reference (sources/models.cpp:24-57 with the conserving
    bond charges {0,-2,2,0,0}), un-coalesced: 5 sections of size 1 on each bond."""
    phys = SPIN_HALF
    sh = shape([HEIS_MPO_BOND, phys, conj_leg(HEIS_MPO_BOND), conj_leg(phys)], (0,))
    dense = np.zeros((5, 2, 5, 2))
    ident = np.eye(2)
    sz = np.diag([1.0, -1.0])
    up = np.array([[0.0, 1.0], [0.0, 0.0]])   # s'=0 (charge +1), s=1 (charge -1)
    dn = up.T
    dense[0, :, 0, :] = ident
    dense[1, :, 0, :] = up
    dense[2, :, 0, :] = dn
    dense[3, :, 0, :] = sz
    dense[4, :, 1, :] = 2 * Jp * dn
    dense[4, :, 2, :] = 2 * Jp * up
    dense[4, :, 3, :] = Jp * sz
    dense[4, :, 4, :] = ident
    out = dict(sh)
    out["blocks"] = {}
    for idx in allowed_indices(sh):
        v = dense[idx]
        if v != 0.0:
            out["blocks"][idx] = np.full((1, 1, 1, 1), v)
    assert sum(float(b.sum() != 0) for b in out["blocks"].values()) == np.count_nonzero(dense)
    return out


def heff_set(n_sec: int, D: int, sigma: float, seed: int = 1234):
    """T3 of SURVEY.md §8d: psi[a,s1,s2,b], left/right environments E[ket bond, MPO bond, bra bond] and the MPO site W
    (the two-site MPO H2 is compute_2sitesHamil(W,W) = tensordot(W,W,{2},{0}).permute({0,1,3,4,2,5}), dmrg.cpp:503-515).
    Leg charges follow the reference's generate_env construction (dmrg.cpp:370-392): L = (beta*, omega*, beta),
    R = (beta, omega, beta*)."""
    rng = np.random.default_rng(seed)
    beta = bond(n_sec, D, sigma, 2)
    psi = rand_like(shape([beta, SPIN_HALF, SPIN_HALF, conj_leg(beta)], (0,)), rng)
    lenv = rand_like(shape([conj_leg(beta), conj_leg(HEIS_MPO_BOND), beta], (0,)), rng)
    renv = rand_like(shape([beta, HEIS_MPO_BOND, conj_leg(beta)], (0,)), rng)
    return psi, heisenberg_W(), lenv, renv


def random_btensor(rng: np.random.Generator, rank: int, max_sec: int = 4, max_size: int = 5, nc: int = 1,
                   fill: float = 0.7, legs=None, sel=None):
    """small random block tensor for property tests: random sections/charges, a random subset of the allowed blocks"""
    if legs is None:
        legs = []
        for _ in range(rank):
            ns = int(rng.integers(1, max_sec + 1))
            sizes = [int(x) for x in rng.integers(1, max_size + 1, ns)]
            charges = [tuple(int(x) for x in rng.integers(-2, 3, nc)) for _ in range(ns)]
            legs.append((sizes, charges))
    if sel is None:  # the flux of a random block, so that at least one block is allowed
        pick = [int(rng.integers(0, len(l[0]))) for l in legs]
        sel = tuple(sum(legs[d][1][i][c] for d, i in enumerate(pick)) for c in range(nc))
    sh = shape(legs, sel)
    out = dict(sh)
    out["blocks"] = {}
    allowed = allowed_indices(sh)
    for n, idx in enumerate(allowed):
        if rng.random() < fill or (n == len(allowed) - 1 and not out["blocks"]):
            dims = tuple(sh["sec_sizes"][d][i] for d, i in enumerate(idx))
            out["blocks"][idx] = rng.standard_normal(dims)
    return out


def stored_bytes(desc) -> int:
    return int(sum(b.size for b in desc["blocks"].values())) * 8


# ---- whole-chain inputs for the two-site DMRG workloads (BASELINE.json configs[0] and configs[2]) ---------------------
def heisenberg_mpo(L: int, Jp: float = 0.25):
    """U(1) Heisenberg S=1/2 open chain as a list of L rank-4 MPO site tensors W[wl, s', wr, s] (same operator as the
    reference's Heisenberg(J=-1, L) after to_bMPO with the conserving bond charges, sources/models.cpp:24-70 and
    SURVEY.md appendix C; not coalesced). Site 0 keeps only the last row of the bulk tensor (left bond = one section of
    charge 0), site L-1 only the first column."""
    bulk = heisenberg_W(Jp)
    one = ([1], [(0,)])

    def cut(left_row=None, right_col=None):
        lb = one if left_row is not None else HEIS_MPO_BOND
        rb = one if right_col is not None else HEIS_MPO_BOND
        sh = shape([lb, SPIN_HALF, conj_leg(rb), conj_leg(SPIN_HALF)], (0,))
        out = dict(sh)
        out["blocks"] = {}
        for (wl_, sp, wr, s), v in bulk["blocks"].items():
            if left_row is not None and wl_ != left_row:
                continue
            if right_col is not None and wr != right_col:
                continue
            out["blocks"][(0 if left_row is not None else wl_, sp, 0 if right_col is not None else wr, s)] = v.copy()
        return out

    assert L >= 2
    return [cut(left_row=4)] + [cut() for _ in range(L - 2)] + [cut(right_col=0)]


def random_mps(L: int, bond: int, target: int, seed: int = 0):
    """Random U(1) bMPS of total charge `target` (sum of the physical charges +-1), bond dimension <= `bond` spread over
    the reachable charge sectors, right-orthonormal on sites 1..L-1 with the orthogonality centre on site 0 (what
    quantit::random_bMPS + move_oc(0) hand to dmrg, sources/MPT.cpp). Site tensors A[l, s, r]: left leg charges q_l,
    physical +-1, right leg -q_r with q_r = q_l + s; selection rule 0; the last right bond is one section of charge
    `target`."""
    rng = np.random.default_rng(seed)
    assert (L + target) % 2 == 0 and abs(target) <= L
    # bond k (between site k-1 and k): charges reachable from the left (k spins) and from the right (L-k spins)
    bonds = []
    for k in range(L + 1):
        qs = [q for q in range(-k, k + 1, 2) if abs(target - q) <= L - k]
        from math import comb
        cap = [min(comb(k, (k + q) // 2), comb(L - k, (L - k + target - q) // 2)) for q in qs]
        per = max(1, bond // max(1, len(qs)))
        sizes = [int(min(c, per)) for c in cap]
        bonds.append((sizes, [(q,) for q in qs]))
    sites = []
    for k in range(L):
        sh = shape([bonds[k], SPIN_HALF, conj_leg(bonds[k + 1])], (0,))
        t = rand_like(sh, rng)
        for key in t["blocks"]:
            t["blocks"][key] = t["blocks"][key] - 0.5
        sites.append(t)
    # right-orthonormalise from the right: per left sector, LQ of [D_l, (s, r)] and push L into the site on the left
    for k in range(L - 1, 0, -1):
        t = sites[k]
        nl = len(t["sec_sizes"][0])
        new_left = list(t["sec_sizes"][0])
        lmats = {}
        for ql in range(nl):
            keys = sorted(key for key in t["blocks"] if key[0] == ql)
            if not keys:
                continue
            M = np.concatenate([t["blocks"][key].reshape(t["blocks"][key].shape[0], -1) for key in keys], axis=1)
            qmat, rmat = np.linalg.qr(M.T)  # M = rmat^T qmat^T
            r = qmat.shape[1]
            assert r == M.shape[0], "bond sector larger than what the right side can support"
            lmats[ql] = rmat.T
            pos = 0
            for key in keys:
                blk = t["blocks"][key]
                w = blk.shape[1] * blk.shape[2]
                t["blocks"][key] = np.ascontiguousarray(qmat.T[:, pos:pos + w]).reshape(blk.shape)
                pos += w
        left = sites[k - 1]
        for key in list(left["blocks"]):
            if key[2] in lmats:
                left["blocks"][key] = np.ascontiguousarray(left["blocks"][key] @ lmats[key[2]])
        t["sec_sizes"][0] = new_left
    nrm = np.sqrt(sum(float((b * b).sum()) for b in sites[0]["blocks"].values()))
    for key in sites[0]["blocks"]:
        sites[0]["blocks"][key] = sites[0]["blocks"][key] / nrm
    return sites


# ---- T4: Hubbard-profile two-site tensor, U(1) x U(1) (charge N, 2 Sz) ------------------------------
HUBBARD_SITE = ([1, 1, 1, 1], [(0, 0), (1, 1), (1, -1), (2, 0)])  # python_binding/exemples/dmrg.py:41


def bond2d(D: int, n_filling: int, sigma_n: float = 1.6, sigma_s: float = 2.2, min_weight: float = 1e-2):
    """bond leg with a 2-D Gaussian profile over (N, 2 Sz): sectors (n, s) with n + s even around (n_filling, 0), sizes
    proportional to exp(-(n - n0)^2 / 2 sigma_n^2 - s^2 / 2 sigma_s^2), total D (the remainder goes to the centre)."""
    secs = []
    for n in range(n_filling - 8, n_filling + 9):
        for s in range(-8, 9):
            if (n + s) % 2:
                continue
            w = float(np.exp(-((n - n_filling) ** 2) / (2 * sigma_n ** 2) - (s ** 2) / (2 * sigma_s ** 2)))
            if w >= min_weight:
                secs.append(((n, s), w))
    tot = sum(w for _, w in secs)
    sizes = [max(1, int(round(D * w / tot))) for _, w in secs]
    centre = max(range(len(secs)), key=lambda i: secs[i][1])
    sizes[centre] += D - sum(sizes)
    return (sizes, [q for q, _ in secs])


def hubbard_theta(D: int, rng: np.random.Generator, n_filling: int = 10):
    """theta = rand_like(shape(beta, site, site, beta'*)) with the Hubbard physical leg, selection rule (0, 0)"""
    beta = bond2d(D, n_filling)
    right = bond2d(D, n_filling + 2)  # the two sites add between 0 and 4 particles: centre the right bond on +2
    return rand_like(shape([beta, HUBBARD_SITE, HUBBARD_SITE, conj_leg(right)], (0, 0)), rng)


# ---- Heisenberg model on an arbitrary bond list as a U(1) MPO (BASELINE.json configs[3]: width-6 cylinder) -----------------
def heisenberg_bonds_mpo(L: int, bonds, Jp: float = 0.25):
    """H = sum_{(i,j) in bonds} Jp (2 S+_i S-_j + 2 S-_i S+_j + Sz_i Sz_j) in the convention of heisenberg_W (Pauli-like
    matrices sz = diag(1,-1), up, dn; the nearest-neighbour chain gives heisenberg_mpo). Finite-state-machine MPO on the
    1-D ordering 0..L-1: MPO bond k (between sites k-1 and k) carries the states `done` (charge 0), one state per pending
    half-bond (i, a) with i < k <= j_max(i), a in {up: charge +2, dn: -2, sz: 0}, and `start` (charge 0); a pending
    operator is shared by every bond (i, j) that starts at i. Every state is a section of size 1 (the reference's
    bMPO::coalesce / qtb_coalesce merges equal charges). Returns L site tensors W[wl, s', wr, s] like heisenberg_mpo."""
    sz = np.diag([1.0, -1.0])
    up = np.array([[0.0, 1.0], [0.0, 0.0]])
    dn = up.T
    ident = np.eye(2)
    ops = {"up": up, "dn": dn, "sz": sz}
    partner = {"up": (dn, 2 * Jp), "dn": (up, 2 * Jp), "sz": (sz, Jp)}
    charge = {"up": 2, "dn": -2, "sz": 0}
    ends = {}
    for (i, j) in bonds:
        i, j = (i, j) if i < j else (j, i)
        assert 0 <= i < j < L
        ends.setdefault(i, set()).add(j)

    def states(k):  # MPO bond k
        if k == 0:
            return [("start",)]
        if k == L:
            return [("done",)]
        st = [("done",)]
        for i in sorted(ends):
            if i < k <= max(ends[i]):
                st += [(i, a) for a in ("up", "dn", "sz")]
        return st + [("start",)]

    def q(state):
        return (0,) if len(state) == 1 else (charge[state[1]],)

    sites = []
    for k in range(L):
        sl, sr = states(k), states(k + 1)
        lb = ([1] * len(sl), [q(t) for t in sl])
        rb = ([1] * len(sr), [q(t) for t in sr])
        sh = shape([lb, SPIN_HALF, conj_leg(rb), conj_leg(SPIN_HALF)], (0,))
        dense = np.zeros((len(sl), 2, len(sr), 2))
        ri = {t: n for n, t in enumerate(sr)}
        for a_, t in enumerate(sl):
            if t == ("done",):
                dense[a_, :, ri[("done",)], :] += ident
            elif t == ("start",):
                if ("start",) in ri:
                    dense[a_, :, ri[("start",)], :] += ident
                for a in ("up", "dn", "sz"):
                    if (k, a) in ri:
                        dense[a_, :, ri[(k, a)], :] += ops[a]
            else:
                i, a = t
                if k in ends[i]:
                    op, c = partner[a]
                    dense[a_, :, ri[("done",)], :] += c * op
                if t in ri:
                    dense[a_, :, ri[t], :] += ident
        out = dict(sh)
        out["blocks"] = {}
        for idx in allowed_indices(sh):
            if dense[idx] != 0.0:
                out["blocks"][idx] = np.full((1, 1, 1, 1), dense[idx])
        assert len(out["blocks"]) == np.count_nonzero(dense), "an MPO element violates the U(1) selection rule"
        sites.append(out)
    return sites


def cylinder_bonds(Lx: int, Ly: int):
    """nearest-neighbour bonds of an Lx x Ly square lattice, periodic around the circumference Ly (a cylinder), open along
    Lx; site (x, y) -> x * Ly + y (column-major snake-free ordering, the one of the reference's 4x8 fixture)"""
    b = []
    for x in range(Lx):
        for y in range(Ly):
            s = x * Ly + y
            if Ly > 2 or y == 0:
                b.append((s, x * Ly + (y + 1) % Ly)) if Ly > 1 else None
            if x + 1 < Lx:
                b.append((s, (x + 1) * Ly + y))
    return sorted(set((min(i, j), max(i, j)) for i, j in b if i != j))


def heisenberg_cylinder_mpo(Lx: int, Ly: int = 6, Jp: float = 0.25):
    """BASELINE.json configs[3]: the width-`Ly` Heisenberg cylinder (the reference ships only the 4x8 fixture files)"""
    return heisenberg_bonds_mpo(Lx * Ly, cylinder_bonds(Lx, Ly), Jp)

