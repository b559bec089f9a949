"""Whole two-site DMRG runs sharded by charge sector over N GPUs (torchrun, NCCL) against the same run on one GPU:
energies per sweep and the final state must be IDENTICAL (the sharded engine computes each owned sector exactly as the
single-rank engine does and only sums disjoint contributions). usage: torchrun --nproc-per-node N profiles/sharded_dmrg_driver.py L maxbond sweeps"""
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

L, maxbond, nsw = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = qb.Context(local)


def run():
    H = [qb.BTensor.from_host(**h, ctx=ctx) for h in wl.heisenberg_mpo(L)]
    psi = [qb.BTensor.from_host(**p, ctx=ctx) for p in wl.random_mps(L, 4, L % 2, seed=0)]
    log = {}
    E = qb.dmrg(H, psi, qb.dmrg_options(1e-12, 0.0, maxbond, 4, nsw), oc=0, log=log)
    return E, log, [p.to_host() for p in psi], qb.contract(psi, psi, H), qb.contract(psi, psi)


E1, log1, psi1, c1, n1 = run()
enable_sharding(ctx)
EN, logN, psiN, cN, nN = run()
same_E = log1["energy"] == logN["energy"]
same_psi = all(sorted(a) == sorted(b) and all(np.array_equal(a[k], b[k]) for k in a) for a, b in zip(psi1, psiN))
ok = torch.tensor([1.0 if (same_E and same_psi) else 0.0], device=f"cuda:{local}")
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world} L {L} maxbond {maxbond}: E {E1:.12f} sharded {EN:.12f} energies identical {same_E} final state identical "
          f"{same_psi} (all ranks: {bool(ok.item())}); <H> {cN:.12f} norm {nN:.12f}; seconds/sweep 1 GPU "
          f"{np.mean(log1['seconds'][-2:]):.3f} sharded {np.mean(logN['seconds'][-2:]):.3f}")
dist.barrier()
dist.destroy_process_group()
