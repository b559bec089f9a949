"""GPU tests of the zero-copy boundary INTEGRATION.md is built around: qtb_tensor_adopt on blocks that are separately
allocated CUDA torch tensors (exactly what a quantit::btensor on a CUDA device holds: one torch::Tensor per block), and
results handed back as device pointers + strides that torch wraps without a copy (the from_blob route of the adaptor).
torch appears here as the container only: no torch op computes anything that is checked."""
import ctypes as C

import numpy as np
import pytest

import qtb_oracle as orc
from quantit_b200 import workloads as wl

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def adopt(qb, desc, make_view, ctx=None):
    dev = {}
    for k, v in desc["blocks"].items():
        dev[k] = make_view(torch.from_numpy(np.ascontiguousarray(v)).cuda())
    return qb.BTensor.adopt(desc["sec_sizes"], desc["cvals"], desc["sel"], dev, ctx=ctx), dev


def strided_view(t):
    """the same values seen through a non-contiguous view: embedded in a larger allocation with an offset and a gap in the
    last dim (odd element offsets on purpose: the bulk-copy staging must fall back / shift correctly)"""
    pad = [s + 3 for s in t.shape]
    big = torch.zeros(pad, dtype=torch.float64, device="cuda")
    sl = tuple(slice(1, 1 + s) for s in t.shape)
    big[sl] = t
    return big[sl]


def transposed_view(t):
    if t.dim() < 2:
        return t
    perm = list(range(t.dim()))[::-1]
    return t.permute(perm).contiguous().permute(perm)  # same logical tensor, reversed memory order


@pytest.mark.parametrize("view", [lambda t: t, strided_view, transposed_view])
@pytest.mark.parametrize("n_sec,D,sigma", [(5, 24, 1.0), (9, 96, 1.5)])
def test_tensordot_on_adopted_torch_blocks(engine, view, n_sec, D, sigma):
    qb = engine
    a, b, da, db = wl.tdot_pair(n_sec, D, sigma, seed=17)
    A, keep_a = adopt(qb, a, view)
    B, keep_b = adopt(qb, b, view)
    Cg = A.tensordot(B, da, db)
    want = orc.tensordot(orc.BT(**a), orc.BT(**b), da, db)
    got = Cg.to_host()
    assert list(got) == sorted(want.blocks)
    den = max(float(np.max(np.abs(v))) for v in want.blocks.values())
    for k, v in want.blocks.items():
        assert np.max(np.abs(got[k] - v)) <= 1e-12 * den
    # the result's blocks are views into ONE device arena: torch wraps them in place (no copy), as the adaptor does with
    # torch::from_blob; reading them back through torch must give the same numbers
    idx, dims, strides, ptrs = Cg.block_table()
    qb.default_context().sync()
    for i, d, s, p in zip(idx, dims, strides, ptrs):
        # torch has no public from_blob in Python: build the zero-copy view through the CUDA array interface
        class _Blk:
            __cuda_array_interface__ = {"shape": tuple(d), "typestr": "<f8", "data": (p, False), "version": 3,
                                        "strides": tuple(8 * x for x in s)}
        w = torch.as_tensor(_Blk(), device="cuda")
        assert w.data_ptr() == p
        assert np.max(np.abs(w.cpu().numpy() - want.blocks[i])) <= 1e-12 * den


def test_svd_and_update_on_adopted_blocks(engine):
    qb = engine
    psi, W, L, R = wl.heff_set(7, 48, 1.3, seed=3)
    P, k1 = adopt(qb, psi, strided_view)
    Wb = qb.BTensor.from_host(**W)
    H2 = Wb.tensordot(Wb, [2], [0]).permute([0, 1, 3, 4, 2, 5])
    Lb, k2 = adopt(qb, L, lambda t: t)
    Rb, k3 = adopt(qb, R, transposed_view)
    E, new = qb.two_sites_update(P, H2, Lb, Rb)
    oW = orc.BT(**W)
    oH2 = orc.permute(orc.tensordot(oW, oW, [2], [0]), [0, 1, 3, 4, 2, 5])
    oE, onew = orc.two_sites_update(orc.BT(**psi), oH2, orc.BT(**L), orc.BT(**R))
    assert abs(E - oE) <= 1e-11 * abs(oE)
    got = new.to_host()
    assert sorted(got) == sorted(onew.blocks)
    U, d, V = qb.svd(P, 2, 1e-10, 4, 48)
    oU, od, oV = orc.svd_trunc(orc.BT(**psi), 2, 1e-10, 4, 48)
    ss, cv, sel, _ = d.structure()
    assert ss == od.sec_sizes
    for k, v in od.blocks.items():
        assert np.max(np.abs(d.to_host()[k] - v)) <= 1e-12


def test_adopted_tensor_survives_context_teardown_order(engine):
    """a tensor freed after its context must not touch the dead context"""
    qb = engine
    ctx = qb.Context(0)
    a, b, da, db = wl.tdot_pair(5, 24, 1.0, seed=2)
    A = qb.BTensor.from_host(**a, ctx=ctx)
    T, keep = adopt(qb, a, lambda t: t, ctx=ctx)
    V = A.permute([3, 2, 1, 0])  # a view sharing A's arena
    handles = [A.h, T.h, V.h]
    A.h = T.h = V.h = None  # take the handles away from the python wrappers
    lib = ctx.lib
    ctx.close()  # qtb_ctx_destroy: the arenas of the three tensors are orphaned, their device blocks released
    for h in handles:  # freed AFTER the context: must not touch it
        lib.qtb_tensor_free(h)
