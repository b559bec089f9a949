"""GPU tests of the charge-sector sharding on ONE device: two engine contexts (rank 0 and rank 1 of
a world of 2) run the same calls from two host threads; the allreduce callback of the C ABI is served through host
memory with a barrier. What is checked is the engine's own logic — owner maps, owned tile lists, zero-filled arenas,
where the collectives sit — and its central claim: a sharded run is BIT-IDENTICAL to the single-rank run.
(The NCCL plumbing itself, quantit_b200/sharding.py, needs >= 2 GPUs: profiles/sharded_driver.py under torchrun.)"""
import json
import os
import threading

import numpy as np
import pytest

import qtb_oracle as orc
from quantit_b200 import workloads as wl
from quantit_b200.sharding import _DeviceBuffer

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class HostAllreduce:
    def __init__(self, world):
        self.world = world
        self.bar = threading.Barrier(world, timeout=180)
        self.bufs = [None] * world
        self.calls = [0] * world
        self.bytes = [0] * world

    def bind(self, rank, ctx):
        import torch

        def ar(ptr, n, stream):
            ctx.sync()
            t = torch.as_tensor(_DeviceBuffer(ptr, n), device="cuda:0")
            self.bufs[rank] = t.cpu().numpy().copy()
            self.bar.wait()
            tot = self.bufs[0].copy()
            for r in range(1, self.world):
                tot += self.bufs[r]
            self.bar.wait()
            t.copy_(torch.from_numpy(tot))
            torch.cuda.synchronize()
            self.calls[rank] += 1
            self.bytes[rank] += 8 * n

        ctx.set_sharding(rank, self.world, ar)


def run_ranks(qb, world, fn):
    """fn(ctx, rank) on `world` threads, one sharded context each; returns the per-rank results"""
    har = HostAllreduce(world)
    out, err = [None] * world, [None] * world

    def body(r):
        try:
            ctx = qb.Context(0)
            har.bind(r, ctx)
            out[r] = fn(ctx, r)
            ctx.sync()
        except BaseException as e:  # noqa: BLE001
            err[r] = e
            har.bar.abort()

    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out, har


@pytest.mark.parametrize("world", [2, 3])
def test_heff_env_svd_sharded_bit_identical(engine, world):
    qb = engine
    psi, W, L, R = wl.heff_set(9, 160, 1.6, seed=5)

    def work(ctx, rank):
        bt = lambda d: qb.BTensor.from_host(**d, ctx=ctx)
        Wb = bt(W)
        H2 = Wb.tensordot(Wb, [2], [0]).permute([0, 1, 3, 4, 2, 5])
        p, l, r = bt(psi), bt(L), bt(R)
        f0 = ctx.counters()["gemm_flops"]  # H2 above is a replicated (unsharded) contraction
        phi = qb.hamil2site_times_state(p, H2, l, r)
        E, p2 = qb.two_sites_update(p, H2, l, r)
        U, d, V = qb.svd(p2, 2, 1e-8, 4, 120)
        le = qb.compute_left_env(Wb, U, l)
        re = qb.compute_right_env(Wb, V.conj().permute([2, 0, 1]), r)
        res = {"E": E, "flops": ctx.counters()["gemm_flops"] - f0}
        for name, t in (("phi", phi), ("p2", p2), ("U", U), ("d", d), ("V", V), ("le", le), ("re", re)):
            res[name] = (t.structure(), t.to_host())
        return res

    ref = work(qb.Context(0), 0)
    outs, har = run_ranks(qb, world, work)
    assert min(har.calls) > 0 and len(set(har.calls)) == 1
    flops = [o["flops"] for o in outs]
    # every rank did a share of the contraction work (the owner maps split it), none did all of it
    assert max(flops) < 0.8 * ref["flops"] and sum(flops) == ref["flops"]
    for o in outs:
        assert o["E"] == ref["E"]
        for name in ("phi", "p2", "U", "d", "V", "le", "re"):
            assert o[name][0] == ref[name][0], name
            assert sorted(o[name][1]) == sorted(ref[name][1]), name
            for k, blk in ref[name][1].items():
                assert np.array_equal(o[name][1][k], blk), (name, k)


def test_heff_sharded_bit_identical_across_tile_configurations(engine):
    """at bond dimension 1536 the single-rank plans take 128 x 128 tiles while the plans of a sharded context take 64 x 64
    (a rank's share of the large tiles would leave SMs idle); the per-element sequence of DMMAs is
    the same, so H_eff.psi must still be bit-identical"""
    qb = engine
    psi, W, L, R = wl.heff_set(9, 1536, 1.6, seed=11)

    def work(ctx, rank):
        bt = lambda d: qb.BTensor.from_host(**d, ctx=ctx)
        Wb = bt(W)
        H2 = Wb.tensordot(Wb, [2], [0]).permute([0, 1, 3, 4, 2, 5])
        phi = qb.hamil2site_times_state(bt(psi), H2, bt(L), bt(R))
        return phi.structure(), phi.to_host()

    ref = work(qb.Context(0), 0)
    outs, _ = run_ranks(qb, 2, work)
    for st, blocks in outs:
        assert st == ref[0] and list(blocks) == list(ref[1])
        for k, blk in ref[1].items():
            assert np.array_equal(blocks[k], blk), k


def test_dmrg_sharded_matches_reference_run(engine):
    """whole two-site DMRG (Heisenberg L=8, the committed reference run) on 2 sharded ranks: per-sweep energies equal
    the single-rank engine run bit for bit and the reference's to 1e-10"""
    qb = engine
    d = os.path.join(G, "dmrg_heis8")
    rec = json.load(open(os.path.join(d, "reference_run.json")))
    L = rec["L"]
    Hh = [orc.read_qtbt(os.path.join(d, f"H_{i}.qtbt")) for i in range(L)]
    Ph = [orc.read_qtbt(os.path.join(d, f"psi0_{i}.qtbt")) for i in range(L)]

    def work(ctx, rank):
        eng = lambda t: qb.BTensor.from_host(t.sec_sizes, t.cvals, t.sel, t.blocks, mods=t.mods, ctx=ctx)
        H, psi = [eng(t) for t in Hh], [eng(t) for t in Ph]
        log = {}
        opt = qb.dmrg_options(rec["cutoff"], rec["convergence_criterion"], rec["maximum_bond"], rec["minimum_bond"],
                              rec["maximum_iterations"])
        E = qb.dmrg(H, psi, opt, oc=rec["oc"], log=log)
        return E, log["energy"], log["mid_bond"], [p.to_host() for p in psi]

    ref = work(qb.Context(0), 0)
    outs, _ = run_ranks(qb, 2, work)
    assert np.allclose(ref[1], rec["sweep_energy"], rtol=1e-10, atol=0)
    for o in outs:
        assert o[0] == ref[0] and o[1] == ref[1] and o[2] == ref[2]
        for a, b in zip(o[3], ref[3]):
            assert sorted(a) == sorted(b)
            for k in a:
                assert np.array_equal(a[k], b[k])
