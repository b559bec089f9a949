"""Sharded H_eff.psi / environment update / block SVD over N GPUs (torchrun, NCCL): checks bit-identity against the
unsharded run on the same rank and times the sharded calls (CUDA events, max over ranks). Diagnostics driver.
usage: torchrun --nproc-per-node N profiles/sharded_driver.py n_sec D sigma [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import quantit_b200 as qb
from quantit_b200 import workloads as wl
from quantit_b200.sharding import enable_sharding

n_sec, D, sigma = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = qb.Context(local)
psi, W, L, R = wl.heff_set(n_sec, D, sigma, seed=5)
bt = lambda d: qb.BTensor.from_host(**d, ctx=ctx)
Wb = bt(W)
H2 = Wb.tensordot(Wb, [2], [0]).permute([0, 1, 3, 4, 2, 5])
p, l, r = bt(psi), bt(L), bt(R)


def timed(fn, n):
    ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    fn()
    ctx.sync()
    dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(ext):
        ev[0].record()
    for _ in range(n):
        out = fn()
    with torch.cuda.stream(ext):
        ev[1].record()
    ctx.sync()
    ms = torch.tensor([ev[0].elapsed_time(ev[1]) / n], device=f"cuda:{local}")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), out


# unsharded reference on this rank
c0 = ctx.counters()["gemm_flops"]
ms1, phi1 = timed(lambda: qb.hamil2site_times_state(p, H2, l, r), reps)
flops = (ctx.counters()["gemm_flops"] - c0) // (reps + 1)

ref = phi1.to_host()
enable_sharding(ctx)
c0 = ctx.counters()["gemm_flops"]
msN, phiN = timed(lambda: qb.hamil2site_times_state(p, H2, l, r), reps)
mine = (ctx.counters()["gemm_flops"] - c0) // (reps + 1)
got = phiN.to_host()
same = sorted(got) == sorted(ref) and all(np.array_equal(got[k], ref[k]) for k in ref)
E, p2 = qb.two_sites_update(p, H2, l, r)
msS, (U, d, V) = timed(lambda: qb.svd(p2, 2, 1e-10, 4, D), max(1, reps // 2))
share = torch.tensor([float(mine)], device=f"cuda:{local}")
shares = [torch.zeros_like(share) for _ in range(world)]
dist.all_gather(shares, share)
if rank == 0:
    print(f"world {world} D {D}: H_eff.psi {flops/1e9:.1f} GFLOP unsharded {ms1:.3f} ms ({flops/ms1/1e9:.2f} TFLOP/s) "
          f"sharded {msN:.3f} ms ({flops/msN/1e9:.2f} TFLOP/s aggregate, speedup {ms1/msN:.2f}x, efficiency {ms1/msN/world:.2f}) "
          f"bit-identical {same}; flop shares {[round(float(s.item())/flops, 3) for s in shares]}; "
          f"sharded svd {msS:.1f} ms kept {sum(d.structure()[0][0])}; E {E:.12f}", flush=True)
dist.barrier()
dist.destroy_process_group()
