"""Host plumbing of the charge-sector sharding: one process per GPU, torch.distributed (NCCL over
NVLink / NVSwitch) supplies the single collective the engine needs — an in-place fp64 sum-allreduce of a device buffer —
through the C-ABI callback of include/qtb.h (qtb_ctx_set_sharding). The engine decides WHAT to reduce (the result arena
of a sharded contraction chain, the singular values, the U / V arenas of the block SVD); nothing here touches tensor
data. Every rank must hold the same tensors and make the same engine calls (SPMD)."""
from __future__ import annotations

from typing import Optional

from .engine import Context, default_context


class _DeviceBuffer:
    """zero-copy view of n doubles at a raw device address for torch.as_tensor (__cuda_array_interface__ v2)"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2,
                                         "strides": None}


def enable_sharding(ctx: Optional[Context] = None, group=None, use_engine_nccl: bool = True) -> Context:
    """shard the engine's contractions and SVDs of `ctx` over the ranks of the torch.distributed process group `group`
    (default: the world group, backend nccl). Call after dist.init_process_group; one context per process."""
    import torch
    import torch.distributed as dist

    ctx = ctx or default_context()
    if not dist.is_initialized():
        raise RuntimeError("enable_sharding: torch.distributed is not initialised")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        ctx.set_sharding(0, 1, None)
        return ctx
    dev = torch.device("cuda", ctx.device)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def allreduce(ptr: int, n: int, stream: int) -> None:
        t = torch.as_tensor(_DeviceBuffer(ptr, n), device=dev)
        with torch.cuda.stream(ext):  # the collective is ordered after the engine's kernels and before its next ones
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    # fallback collective (a host callback into torch.distributed) first, then the engine's own NCCL communicator: with
    # it no Python runs on the data path and the sharded contraction chains exchange exactly the owned row ranges
    ctx.set_sharding(rank, world, allreduce)
    ctx._sharding_keepalive = (ext, group)
    if use_engine_nccl and dist.get_backend(group) == "nccl":
        from .engine import nccl_unique_id

        path = _find_libnccl()
        box = [nccl_unique_id(path) if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ctx.init_nccl(rank, world, box[0], path)
        ctx.engine_nccl = True
    return ctx


def _find_libnccl() -> Optional[str]:
    """the libnccl.so.2 torch itself uses (pip wheels ship it under nvidia/nccl/lib); None = let dlopen search"""
    import glob
    import os
    import sys

    for base in sys.path:
        hits = glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2"))
        if hits:
            return hits[0]
    return None
