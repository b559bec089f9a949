"""One rank's share of a sharded H_eff.psi at D=4096, alone on one GPU (the collective replaced by a no-op): what the
compute side of the strong-scaling run costs per rank, without NVLink. usage: python profiles/r2/rank_alone.py world"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import quantit_b200 as qb
from quantit_b200 import workloads as wl
world = int(sys.argv[1])
ctx = qb.Context(0)
psi, W, L, R = wl.heff_set(15, 4096, 1.6, seed=5)
bt = lambda d: qb.BTensor.from_host(**d, ctx=ctx)
Wb = bt(W); H2 = Wb.tensordot(Wb, [2], [0]).permute([0, 1, 3, 4, 2, 5]); p, l, r = bt(psi), bt(L), bt(R)
ext = torch.cuda.ExternalStream(ctx.stream)
def timed(n=10):
    qb.hamil2site_times_state(p, H2, l, r); ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(ext): e0.record()
    for _ in range(n): o = qb.hamil2site_times_state(p, H2, l, r)
    t1 = time.perf_counter()
    with torch.cuda.stream(ext): e1.record()
    ctx.sync()
    return e0.elapsed_time(e1) / n, (t1 - t0) / n * 1e3
print("unsharded: gpu %.3f ms, host issue %.3f ms" % timed())
for rank in range(min(world, 3)):
    ctx.set_sharding(rank, world, lambda ptr, n, stream: None)
    g, h = timed()
    c = ctx.counters()
    print(f"world {world} rank {rank} alone: gpu {g:.3f} ms, host issue {h:.3f} ms per call")
