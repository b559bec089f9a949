"""CPU tests of the host side of the charge-sector sharding: the LPT sector assignment of the
C ABI (qtb_lpt_assign), and — with two gloo ranks — the property the design relies on: every rank fills only the
sections it owns of a zero buffer, a sum-allreduce makes the buffer whole and identical on all ranks, bit for bit."""
import os
import socket

import numpy as np
import pytest


def test_lpt_assign_properties(engine):
    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 5, 15, 41):
            w = rng.random(n) ** 3 * 1e9
            own = engine.engine.lpt_assign(w, world)
            assert len(own) == n and all(0 <= r < world for r in own)
            if n == 0:
                continue
            load = np.zeros(world)
            for s, r in enumerate(own):
                load[r] += w[s]
            # LPT guarantee: makespan <= (4/3 - 1/(3 world)) * optimum, optimum >= max(mean load, heaviest section)
            opt_lb = max(w.sum() / world, w.max())
            assert load.max() <= (4.0 / 3.0) * opt_lb + 1e-9
    # deterministic, ties -> lower section to lower rank
    assert engine.engine.lpt_assign([1.0, 1.0, 1.0, 1.0], 2) == [0, 1, 0, 1]
    # a Gaussian sector profile like the D=4096 Heisenberg bond: the centre sector bounds the balance at 8 ranks
    c = np.exp(-((np.arange(15) - 7) ** 2) / (2 * 1.6 ** 2))
    own = engine.engine.lpt_assign(c ** 2, 2)
    load = [sum(c[s] ** 2 for s in range(15) if own[s] == r) for r in range(2)]
    assert abs(load[0] - load[1]) / sum(load) < 0.05


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, weights, sizes, out_q):
    import torch
    import torch.distributed as dist
    import quantit_b200 as qb

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        own = qb.engine.lpt_assign(weights, world)
        full = np.concatenate([np.random.default_rng(100 + s).standard_normal(n) for s, n in enumerate(sizes)])
        offs = np.concatenate([[0], np.cumsum(sizes)])
        buf = np.zeros_like(full)
        for s in range(len(sizes)):
            if own[s] == rank:
                buf[offs[s]:offs[s + 1]] = full[offs[s]:offs[s + 1]]
        t = torch.from_numpy(buf)
        dist.all_reduce(t)
        out_q.put((rank, own, bool(np.array_equal(t.numpy(), full))))
    finally:
        dist.destroy_process_group()


def test_two_ranks_owned_sections_sum_to_whole(engine):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    sizes = [3, 40, 170, 500, 170, 40, 3]
    weights = [float(n) ** 3 for n in sizes]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, weights, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1], "ranks disagree on the owner map"
    assert set(res[0][1]) == {0, 1}
    assert res[0][2] and res[1][2], "allreduce of owned pieces over zeros is not bit-exact"
