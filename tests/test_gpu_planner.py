"""Device block-pair matching (quantit_b200/csrc/qtb_match.cu) against the host planner and the oracle.

The north star asks for the selection-rule matching of btensor::tensordot (reference btensor.cpp:2057-2108: the
two-pointer merge over "columns") as an on-GPU sort / segmented-match kernel whose output block structure is bit-exact.
One context is forced onto the device kernels, one onto the host sort; the two plans must agree in every count, the
results must be IDENTICAL bit for bit (same plan -> same kernel schedule -> same arithmetic) and match the oracle."""
import numpy as np
import pytest

import qtb_oracle as orc
from quantit_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def both_contexts(qb):
    host, dev = qb.Context(0), qb.Context(0)
    host.set_device_planner(0)
    dev.set_device_planner(1)
    return host, dev


def load(qb, d, ctx):
    return qb.BTensor.from_host(**{k: d[k] for k in ("sec_sizes", "cvals", "sel", "blocks")}, ctx=ctx)


def same_bits(x, y):
    hx, hy = x.to_host(), y.to_host()
    assert list(hx) == list(hy), "block order differs between the host and the device matching"
    assert x.structure() == y.structure()
    for k in hx:
        assert hx[k].shape == hy[k].shape and np.array_equal(hx[k], hy[k]), k


@pytest.mark.parametrize("n_sec,D,sigma", [(5, 24, 1.0), (13, 300, 2.0), (41, 2048, 6.0)])
def test_device_matching_rank4(engine, n_sec, D, sigma):
    qb = engine
    host, dev = both_contexts(qb)
    a, b, da, db = wl.tdot_pair(n_sec, D, sigma, seed=7)
    Ch = load(qb, a, host).tensordot(load(qb, b, host), da, db)
    Ad, Bd = load(qb, a, dev), load(qb, b, dev)
    Cd = Ad.tensordot(Bd, da, db)
    assert dev.device_matches() == 1 and host.device_matches() == 0
    assert Ad.tensordot_info(Bd, da, db) == load(qb, a, host).tensordot_info(load(qb, b, host), da, db)
    same_bits(Ch, Cd)
    want = orc.tensordot(orc.BT(**{k: a[k] for k in ("sec_sizes", "cvals", "sel", "blocks")}),
                         orc.BT(**{k: b[k] for k in ("sec_sizes", "cvals", "sel", "blocks")}), da, db)
    got = Cd.to_host()
    assert list(got) == sorted(want.blocks)
    den = max(float(np.max(np.abs(v))) for v in want.blocks.values() if v.size)
    for k, v in want.blocks.items():
        assert float(np.max(np.abs(got[k] - v))) <= 1e-12 * den


@pytest.mark.parametrize("seed", range(10))
def test_device_matching_random_structures(engine, seed):
    """random ranks / sections / charges / missing blocks, multi-index contractions (also rank-0 outputs, k = 0)"""
    qb = engine
    host, dev = both_contexts(qb)
    rng = np.random.default_rng(500 + seed)
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
    Ch = load(qb, a, host).tensordot(load(qb, b, host), dims_a, dims_b)
    Cd = load(qb, a, dev).tensordot(load(qb, b, dev), dims_a, dims_b)
    same_bits(Ch, Cd)


def test_device_matching_hubbard_theta(engine):
    """U(1)xU(1) two-site tensor against itself over both bonds: hundreds of blocks, two charge components"""
    qb = engine
    host, dev = both_contexts(qb)
    rng = np.random.default_rng(3)
    th = wl.hubbard_theta(192, rng)
    c = {k: th[k] for k in ("sec_sizes", "cvals", "sel", "blocks")}
    thc = dict(c, cvals=[[tuple(-x for x in q) for q in leg] for leg in c["cvals"]], sel=tuple(-x for x in c["sel"]))
    Ch = load(qb, c, host).tensordot(load(qb, thc, host), [0, 3], [0, 3])
    Cd = load(qb, c, dev).tensordot(load(qb, thc, dev), [0, 3], [0, 3])
    assert dev.device_matches() == 1
    same_bits(Ch, Cd)
