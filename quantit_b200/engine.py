"""ctypes binding of libqtb.so (include/qtb.h) and the host-side mirror of the reference's btensor interface.

The class/method names follow the reference's public API for the hot path (reference include/blockTensor/btensor.h,
include/blockTensor/LinearAlgebra.h, include/dmrg.h) so that parity tests read like the reference's own tests:
`BTensor.tensordot`, `.permute`, `.conj`, `svd(t, split, tol, min, max, pow)`, `hamil2site_times_state`, ...

There is NO CPU fallback: if libqtb.so is missing or no CUDA device is usable, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqtb.so")

i64 = C.c_int64
p_i64 = C.POINTER(C.c_int64)
p_f64 = C.POINTER(C.c_double)
vp = C.c_void_p


class QtbError(RuntimeError):
    """Base class; subclasses mirror the exception types the reference throws."""

    status = 0


class InvalidArgument(QtbError, ValueError):  # std::invalid_argument
    status = 1


class OutOfRange(QtbError, IndexError):  # std::out_of_range
    status = 2


class DomainError(QtbError):  # std::domain_error
    status = 3


class LogicError(QtbError):  # std::logic_error
    status = 4


class EngineRuntimeError(QtbError):  # std::runtime_error
    status = 5


class CheckError(QtbError):  # c10::Error raised by TORCH_CHECK in the reference
    status = 6


class CudaError(QtbError):
    status = 7


class NoDeviceError(QtbError):
    status = 8


_STATUS = {c.status: c for c in (InvalidArgument, OutOfRange, DomainError, LogicError, EngineRuntimeError, CheckError,
                                 CudaError, NoDeviceError)}

# every symbol include/qtb.h declares: (name, restype, argtypes)
_SIGS = [
    ("qtb_last_error", C.c_char_p, []),
    ("qtb_version", C.c_char_p, []),
    ("qtb_ctx_create", C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    ("qtb_ctx_destroy", None, [vp]),
    ("qtb_ctx_sync", C.c_int, [vp]),
    ("qtb_ctx_trim", C.c_int, [vp]),
    ("qtb_ctx_set_device_planner", C.c_int, [vp, C.c_int]),
    ("qtb_ctx_device_matches", C.c_int, [vp, p_i64]),
    ("qtb_ctx_stream", vp, [vp]),
    ("qtb_ctx_counters", C.c_int, [vp, p_i64]),
    ("qtb_ctx_set_sharding", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    ("qtb_nccl_unique_id", C.c_int, [C.c_char_p, C.c_char_p]),
    ("qtb_ctx_init_nccl", C.c_int, [vp, C.c_int, C.c_int, C.c_char_p, C.c_char_p]),
    ("qtb_lpt_assign", C.c_int, [i64, p_f64, C.c_int, C.POINTER(C.c_int32)]),
    ("qtb_tensor_create", C.c_int, [vp, i64, i64, p_i64, p_i64, p_i64, p_i64, p_i64, i64, p_i64, p_f64, C.POINTER(vp)]),
    ("qtb_tensor_adopt", C.c_int, [vp, i64, i64, p_i64, p_i64, p_i64, p_i64, p_i64, i64, p_i64, C.POINTER(vp), p_i64,
                                   C.POINTER(vp)]),
    ("qtb_tensor_free", None, [vp]),
    ("qtb_tensor_rank", i64, [vp]),
    ("qtb_tensor_nc", i64, [vp]),
    ("qtb_tensor_nblocks", i64, [vp]),
    ("qtb_tensor_total_sections", i64, [vp]),
    ("qtb_tensor_numel", i64, [vp]),
    ("qtb_tensor_structure", C.c_int, [vp, p_i64, p_i64, p_i64, p_i64, p_i64]),
    ("qtb_tensor_blocks", C.c_int, [vp, p_i64, p_i64, p_i64, C.POINTER(vp)]),
    ("qtb_tensor_download", C.c_int, [vp, vp, p_f64]),
    ("qtb_permute", C.c_int, [vp, vp, p_i64, C.POINTER(vp)]),
    ("qtb_conj", C.c_int, [vp, vp, C.POINTER(vp)]),
    ("qtb_tensordot", C.c_int, [vp, vp, vp, i64, p_i64, p_i64, C.POINTER(vp)]),
    ("qtb_tensordot_plan_info", C.c_int, [vp, vp, vp, i64, p_i64, p_i64, p_i64, p_i64, p_i64]),
    ("qtb_tensordot_into", C.c_int, [vp, vp, vp, i64, p_i64, p_i64, vp]),
    ("qtb_tensordot_host", C.c_int, [vp, i64, p_i64,
                                     i64, p_i64, p_i64, p_i64, p_i64, i64, p_i64, p_f64,
                                     i64, p_i64, p_i64, p_i64, p_i64, i64, p_i64, p_f64,
                                     i64, p_i64, p_i64, p_i64, p_i64, p_i64, p_f64]),
    ("qtb_axpby", C.c_int, [vp, C.c_double, vp, C.c_double, vp, C.POINTER(vp)]),
    ("qtb_add", C.c_int, [vp, vp, vp, C.c_double, C.POINTER(vp)]),
    ("qtb_dot", C.c_int, [vp, vp, vp, p_f64]),
    ("qtb_scale_", C.c_int, [vp, vp, C.c_double]),
    ("qtb_mul_lastdim", C.c_int, [vp, vp, vp, C.POINTER(vp)]),
    ("qtb_svd", C.c_int, [vp, vp, i64, C.c_int, C.c_double, i64, i64, C.c_double, C.POINTER(vp), C.POINTER(vp),
                          C.POINTER(vp)]),
    ("qtb_truncate", C.c_int, [vp, vp, vp, vp, i64, i64, C.c_double, C.c_double, C.POINTER(vp), C.POINTER(vp),
                               C.POINTER(vp)]),
    ("qtb_eigh", C.c_int, [vp, vp, i64, C.c_int, C.c_double, i64, i64, C.c_double, C.POINTER(vp), C.POINTER(vp)]),
    ("qtb_reshape", C.c_int, [vp, vp, i64, p_i64, C.POINTER(vp)]),
    ("qtb_reshape_as", C.c_int, [vp, vp, vp, C.c_int, C.POINTER(vp)]),
    ("qtb_tensorgdot", C.c_int, [vp, vp, vp, vp, i64, p_i64, p_i64, C.c_double, C.c_double, C.POINTER(vp)]),
    ("qtb_tensor_save", C.c_int, [vp, vp, C.c_char_p]),
    ("qtb_tensor_load", C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
    ("qtb_heff_apply", C.c_int, [vp, vp, vp, vp, vp, C.POINTER(vp)]),
    ("qtb_env_left", C.c_int, [vp, vp, vp, vp, C.POINTER(vp)]),
    ("qtb_env_right", C.c_int, [vp, vp, vp, vp, C.POINTER(vp)]),
    ("qtb_two_sites_update", C.c_int, [vp, vp, vp, vp, vp, p_f64, C.POINTER(vp)]),
    ("qtb_dmrg", C.c_int, [vp, i64, C.POINTER(vp), C.POINTER(vp), p_i64, vp, p_f64, p_i64, p_f64, p_f64, p_i64]),
    ("qtb_dmrg_logged", C.c_int, [vp, i64, C.POINTER(vp), C.POINTER(vp), p_i64, vp, p_f64, p_i64, C.c_void_p, C.c_void_p]),
    ("qtb_contract", C.c_int, [vp, i64, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), p_f64]),
    ("qtb_move_oc", C.c_int, [vp, i64, C.POINTER(vp), p_i64, i64]),
    ("qtb_coalesce", C.c_int, [vp, i64, C.POINTER(vp), C.c_double]),
]
EXPORTED_SYMBOLS = [s[0] for s in _SIGS]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


def nccl_unique_id(libnccl_path: Optional[str] = None) -> bytes:
    """ncclGetUniqueId through the engine's dlopen'ed libnccl (rank 0 calls it, the bytes are broadcast by the host)"""
    buf = C.create_string_buffer(128)
    _check(load_library().qtb_nccl_unique_id(libnccl_path.encode() if libnccl_path else None, buf))
    return buf.raw


def lpt_assign(weights: Sequence[float], world: int) -> List[int]:
    """longest-processing-time-first assignment of weighted charge sectors to ranks (qtb_lpt_assign, pure host)"""
    w = np.ascontiguousarray(np.asarray(weights, dtype=np.float64).reshape(-1))
    out = (C.c_int32 * max(len(w), 1))()
    _check(load_library().qtb_lpt_assign(len(w), _pf(w), int(world), out))
    return [int(out[i]) for i in range(len(w))]


class DmrgOptionsC(C.Structure):
    _fields_ = [("cutoff", C.c_double), ("convergence_criterion", C.c_double), ("maximum_bond", C.c_int64),
                ("minimum_bond", C.c_int64), ("maximum_iterations", C.c_int64)]

_lib = None


def load_library() -> C.CDLL:
    """dlopen libqtb.so (built in-tree by __graft_entry__.build()). Loading needs no GPU; computing does."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                              "quantit_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, res, args in _SIGS:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(st: int) -> None:
    if st != 0:
        msg = load_library().qtb_last_error().decode()
        raise _STATUS.get(st, QtbError)(msg)


def _arr(x, dtype=np.int64):
    a = np.ascontiguousarray(np.asarray(x, dtype=dtype).reshape(-1))
    return a


def _pi(a: np.ndarray):
    return a.ctypes.data_as(p_i64)


def _pf(a: np.ndarray):
    return a.ctypes.data_as(p_f64)


class Context:
    """One engine context = one CUDA device + stream + memory pool + plan cache (qtb_ctx)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = load_library()
        h = vp()
        _check(self.lib.qtb_ctx_create(int(device), vp(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = device

    def sync(self) -> None:
        _check(self.lib.qtb_ctx_sync(self.h))

    def trim_cache(self) -> None:
        """return the engine's cached free device blocks to the driver (qtb_ctx_trim)"""
        _check(self.lib.qtb_ctx_trim(self.h))

    def set_device_planner(self, mode: int) -> None:
        """where the block-pair matching of a contraction runs: 0 host, 1 device kernels, -1 by size
        (qtb_ctx_set_device_planner; reference btensor.cpp:2057-2108)"""
        _check(self.lib.qtb_ctx_set_device_planner(self.h, int(mode)))

    def device_matches(self) -> int:
        out = (C.c_int64 * 1)()
        _check(self.lib.qtb_ctx_device_matches(self.h, out))
        return int(out[0])

    @property
    def stream(self) -> int:
        return int(self.lib.qtb_ctx_stream(self.h) or 0)

    def counters(self) -> Dict[str, int]:
        out = (C.c_int64 * 8)()
        _check(self.lib.qtb_ctx_counters(self.h, out))
        names = ["kernel_launches", "gemm_launches", "plans_built", "plan_cache_hits", "h2d_bytes", "d2h_bytes",
                 "gemm_flops", "device_bytes"]
        return dict(zip(names, [int(v) for v in out]))

    def set_sharding(self, rank: int, world: int, allreduce=None) -> None:
        """charge-sector sharding over `world` ranks (include/qtb.h, qtb_ctx_set_sharding). `allreduce(ptr, n, stream)`
        must enqueue an in-place fp64 sum-allreduce of n doubles at device address ptr, ordered after the prior work
        of CUDA stream `stream` (quantit_b200/sharding.py supplies the torch.distributed / NCCL one)."""
        if allreduce is None:
            self._ar_cb = None
            _check(self.lib.qtb_ctx_set_sharding(self.h, int(rank), int(world), None, None))
            self.rank, self.world = int(rank), int(world)
            return

        def _cb(user, ptr, n, stream):
            try:
                allreduce(int(ptr or 0), int(n), int(stream or 0))
                return 0
            except Exception:  # no exception may cross the C ABI
                import traceback
                traceback.print_exc()
                return 1

        self._ar_cb = ALLREDUCE_FN(_cb)  # keep the trampoline alive as long as the context uses it
        _check(self.lib.qtb_ctx_set_sharding(self.h, int(rank), int(world), C.cast(self._ar_cb, C.c_void_p), None))
        self.rank, self.world = int(rank), int(world)

    def init_nccl(self, rank: int, world: int, unique_id: bytes, libnccl_path: Optional[str] = None) -> None:
        """collective: creates the engine's own NCCL communicator from the 128-byte id of nccl_unique_id()"""
        assert len(unique_id) == 128
        _check(self.lib.qtb_ctx_init_nccl(self.h, int(rank), int(world), unique_id,
                                          libnccl_path.encode() if libnccl_path else None))
        self.rank, self.world = int(rank), int(world)

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.qtb_ctx_destroy(self.h)
            self.h = None


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


class BTensor:
    """Device-resident block tensor (mirror of quantit::btensor for the hot path)."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self.h = handle

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.qtb_tensor_free(self.h)
        except Exception:
            pass
        self.h = None

    # ---- construction ---------------------------------------------------------------------------------------------
    @classmethod
    def from_host(cls, sec_sizes: Sequence[Sequence[int]], cvals: Sequence[Sequence[Tuple[int, ...]]],
                  sel: Tuple[int, ...], blocks: Dict[Tuple[int, ...], np.ndarray], mods: Optional[Sequence[int]] = None,
                  ctx: Optional[Context] = None) -> "BTensor":
        ctx = ctx or default_context()
        rank = len(sec_sizes)
        nc = len(sel)
        nsec = _arr([len(s) for s in sec_sizes])
        ss = _arr([x for s in sec_sizes for x in s])
        cv = _arr([c for cs in cvals for q in cs for c in q])
        sl = _arr(sel)
        md = _arr(mods) if mods is not None else None
        keys = list(blocks.keys())
        idx = _arr([i for k in keys for i in k])
        for k in keys:
            want = tuple(sec_sizes[d][i] for d, i in enumerate(k))
            if tuple(blocks[k].shape) != want:
                raise InvalidArgument(f"block {k} has shape {blocks[k].shape}, sections say {want}")
        data = (np.concatenate([np.ascontiguousarray(blocks[k], dtype=np.float64).reshape(-1) for k in keys])
                if keys else np.zeros(0))
        data = np.ascontiguousarray(data, dtype=np.float64)
        h = vp()
        _check(ctx.lib.qtb_tensor_create(ctx.h, rank, nc, _pi(md) if md is not None else None, _pi(nsec), _pi(ss),
                                         _pi(cv), _pi(sl), len(keys), _pi(idx), _pf(data) if data.size else None,
                                         C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def adopt(cls, sec_sizes: Sequence[Sequence[int]], cvals: Sequence[Sequence[Tuple[int, ...]]], sel: Tuple[int, ...],
              blocks: Dict[Tuple[int, ...], object], mods: Optional[Sequence[int]] = None,
              ctx: Optional[Context] = None) -> "BTensor":
        """Zero-copy wrap of blocks that already live on the device (qtb_tensor_adopt): every value of `blocks` is a
        float64 CUDA array exposing data_ptr() / stride() / shape (a torch.Tensor, possibly a strided view — what a
        quantit::btensor holds per block). The arrays are kept alive by the returned object; the memory is not owned."""
        ctx = ctx or default_context()
        rank, nc = len(sec_sizes), len(sel)
        nsec = _arr([len(s) for s in sec_sizes])
        ss = _arr([x for s in sec_sizes for x in s])
        cv = _arr([c for cs in cvals for q in cs for c in q])
        sl = _arr(sel)
        md = _arr(mods) if mods is not None else None
        keys = list(blocks.keys())
        for k in keys:
            want = tuple(sec_sizes[d][i] for d, i in enumerate(k))
            if tuple(blocks[k].shape) != want:
                raise InvalidArgument(f"block {k} has shape {tuple(blocks[k].shape)}, sections say {want}")
        idx = _arr([i for k in keys for i in k])
        strides = _arr([int(s) for k in keys for s in blocks[k].stride()])
        ptrs = (vp * max(len(keys), 1))(*[int(blocks[k].data_ptr()) for k in keys])
        h = vp()
        _check(ctx.lib.qtb_tensor_adopt(ctx.h, rank, nc, _pi(md) if md is not None else None, _pi(nsec), _pi(ss), _pi(cv),
                                        _pi(sl), len(keys), _pi(idx), ptrs, _pi(strides), C.byref(h)))
        out = cls(ctx, h)
        out._keepalive = [blocks[k] for k in keys]
        return out

    # ---- structure queries ----------------------------------------------------------------------------------------
    def dim(self) -> int:
        return int(self.ctx.lib.qtb_tensor_rank(self.h))

    @property
    def rank(self) -> int:
        return self.dim()

    @property
    def nblocks(self) -> int:
        return int(self.ctx.lib.qtb_tensor_nblocks(self.h))

    def numel(self) -> int:
        return int(self.ctx.lib.qtb_tensor_numel(self.h))

    def structure(self):
        """(sec_sizes per dim, cvals per dim, sel, mods)"""
        lib = self.ctx.lib
        r, nc = self.dim(), int(lib.qtb_tensor_nc(self.h))
        tot = int(lib.qtb_tensor_total_sections(self.h))
        nsec = np.zeros(max(r, 1), np.int64)
        ss = np.zeros(max(tot, 1), np.int64)
        cv = np.zeros(max(tot * nc, 1), np.int64)
        sl = np.zeros(nc, np.int64)
        md = np.zeros(nc, np.int64)
        _check(lib.qtb_tensor_structure(self.h, _pi(nsec), _pi(ss), _pi(cv), _pi(sl), _pi(md)))
        sec_sizes, cvals, k = [], [], 0
        for d in range(r):
            n = int(nsec[d])
            sec_sizes.append([int(x) for x in ss[k:k + n]])
            cvals.append([tuple(int(x) for x in cv[(k + s) * nc:(k + s + 1) * nc]) for s in range(n)])
            k += n
        return sec_sizes, cvals, tuple(int(x) for x in sl), tuple(int(x) for x in md)

    def section_numbers(self) -> List[int]:
        return [len(s) for s in self.structure()[0]]

    def sizes(self) -> List[int]:
        return [sum(s) for s in self.structure()[0]]

    @property
    def selection_rule(self) -> Tuple[int, ...]:
        return self.structure()[2]

    def block_table(self):
        """(index [nb][rank], dims [nb][rank], strides [nb][rank], device pointers [nb])"""
        r, nb = self.dim(), self.nblocks
        idx = np.zeros(max(nb * r, 1), np.int64)
        dims = np.zeros(max(nb * r, 1), np.int64)
        st = np.zeros(max(nb * r, 1), np.int64)
        ptrs = (vp * max(nb, 1))()
        _check(self.ctx.lib.qtb_tensor_blocks(self.h, _pi(idx), _pi(dims), _pi(st), ptrs))
        rs = lambda a: [tuple(int(x) for x in a[b * r:(b + 1) * r]) for b in range(nb)]
        return rs(idx), rs(dims), rs(st), [int(ptrs[b] or 0) for b in range(nb)]

    def block_indices(self) -> List[Tuple[int, ...]]:
        return self.block_table()[0]

    def to_host(self) -> Dict[Tuple[int, ...], np.ndarray]:
        """blocks as C-contiguous numpy arrays (one device->host copy of the packed arena)"""
        idx, dims, _, _ = self.block_table()
        n = self.numel()
        buf = np.zeros(max(n, 1), np.float64)
        _check(self.ctx.lib.qtb_tensor_download(self.ctx.h, self.h, _pf(buf)))
        out, pos = {}, 0
        for i, d in zip(idx, dims):
            m = int(np.prod(d)) if len(d) else 1
            out[i] = buf[pos:pos + m].reshape(d).copy()
            pos += m
        return out

    def item(self) -> float:
        blocks = self.to_host()
        if not blocks:
            return 0.0
        return float(next(iter(blocks.values())).reshape(-1)[0])

    # ---- ops (reference btensor.h:623 tensordot, btensor.cpp:1754 permute, :2156 conj) -------------------------------
    def tensordot(self, other: "BTensor", dims_self: Sequence[int], dims_other: Sequence[int]) -> "BTensor":
        if len(dims_self) != len(dims_other):
            raise CheckError("both dimension lists should have the same length.")
        da, db = _arr(dims_self), _arr(dims_other)
        h = vp()
        _check(self.ctx.lib.qtb_tensordot(self.ctx.h, self.h, other.h, len(da), _pi(da), _pi(db), C.byref(h)))
        return BTensor(self.ctx, h)

    def tensordot_(self, other: "BTensor", dims_self, dims_other, out: "BTensor") -> "BTensor":
        da, db = _arr(dims_self), _arr(dims_other)
        _check(self.ctx.lib.qtb_tensordot_into(self.ctx.h, self.h, other.h, len(da), _pi(da), _pi(db), out.h))
        return out

    def tensordot_info(self, other: "BTensor", dims_self, dims_other) -> Dict[str, int]:
        da, db = _arr(dims_self), _arr(dims_other)
        a, b, c = i64(), i64(), i64()
        _check(self.ctx.lib.qtb_tensordot_plan_info(self.ctx.h, self.h, other.h, len(da), _pi(da), _pi(db),
                                                    C.byref(a), C.byref(b), C.byref(c)))
        return {"out_blocks": a.value, "pairs": b.value, "flops": c.value}

    def permute(self, perm: Sequence[int]) -> "BTensor":
        p = _arr(perm)
        if len(p) != self.dim():
            raise InvalidArgument("permutation length differs from the tensor rank")
        h = vp()
        _check(self.ctx.lib.qtb_permute(self.ctx.h, self.h, _pi(p), C.byref(h)))
        return BTensor(self.ctx, h)

    def conj(self) -> "BTensor":
        h = vp()
        _check(self.ctx.lib.qtb_conj(self.ctx.h, self.h, C.byref(h)))
        return BTensor(self.ctx, h)

    def add(self, other: "BTensor", alpha: float = 1.0) -> "BTensor":
        """reference btensor::add(other, alpha), btensor.cpp:2666 (including its merge behaviour, see qtb.h)"""
        h = vp()
        _check(self.ctx.lib.qtb_add(self.ctx.h, self.h, other.h, float(alpha), C.byref(h)))
        return BTensor(self.ctx, h)

    def axpby(self, alpha: float, other: "BTensor", beta: float) -> "BTensor":
        h = vp()
        _check(self.ctx.lib.qtb_axpby(self.ctx.h, float(alpha), self.h, float(beta), other.h, C.byref(h)))
        return BTensor(self.ctx, h)

    def dot(self, other: "BTensor") -> float:
        out = C.c_double()
        _check(self.ctx.lib.qtb_dot(self.ctx.h, self.h, other.h, C.byref(out)))
        return out.value

    def mul_(self, s: float) -> "BTensor":
        _check(self.ctx.lib.qtb_scale_(self.ctx.h, self.h, float(s)))
        return self

    def mul_lastdim(self, d: "BTensor") -> "BTensor":
        h = vp()
        _check(self.ctx.lib.qtb_mul_lastdim(self.ctx.h, self.h, d.h, C.byref(h)))
        return BTensor(self.ctx, h)

    def reshape(self, index_groups: Sequence[int]) -> "BTensor":
        """reference btensor::reshape(index_groups), btensor.cpp:2986: merges the groups of consecutive dims delimited by
        `index_groups` (e.g. [split] -> a rank-2 tensor)"""
        g = _arr(index_groups)
        h = vp()
        _check(self.ctx.lib.qtb_reshape(self.ctx.h, self.h, len(g), _pi(g) if len(g) else None, C.byref(h)))
        return BTensor(self.ctx, h)

    def reshape_as(self, like: "BTensor", overwrite_c_vals: bool = False) -> "BTensor":
        """reference btensor::reshape_as<mode>(other), btensor.cpp:3026"""
        h = vp()
        _check(self.ctx.lib.qtb_reshape_as(self.ctx.h, self.h, like.h, 1 if overwrite_c_vals else 0, C.byref(h)))
        return BTensor(self.ctx, h)

    def tensorgdot(self, mul1: "BTensor", mul2: "BTensor", dims1, dims2, beta: float = 1.0, alpha: float = 1.0) -> "BTensor":
        """alpha * self + beta * mul1.mul2 — btensor::tensorgdot as declared at reference btensor.h:624 (semantics of the
        dense include/tensorgdot.h:22)"""
        da, db = _arr(dims1), _arr(dims2)
        h = vp()
        _check(self.ctx.lib.qtb_tensorgdot(self.ctx.h, self.h, mul1.h, mul2.h, len(da), _pi(da), _pi(db), float(beta),
                                           float(alpha), C.byref(h)))
        return BTensor(self.ctx, h)

    def save(self, path: str) -> None:
        _check(self.ctx.lib.qtb_tensor_save(self.ctx.h, self.h, str(path).encode()))

    @classmethod
    def load(cls, path: str, ctx: Optional[Context] = None) -> "BTensor":
        ctx = ctx or default_context()
        h = vp()
        _check(ctx.lib.qtb_tensor_load(ctx.h, str(path).encode(), C.byref(h)))
        return cls(ctx, h)


def tensordot(a: BTensor, b: BTensor, dims_a, dims_b) -> BTensor:
    """free function, reference btensor.h:1230"""
    return a.tensordot(b, dims_a, dims_b)


def svd(a: BTensor, split: int, tol: Optional[float] = None, min_size: int = 1, max_size: Optional[int] = None,
        pow: float = 2.0):
    """reference quantit::svd(btensor, split) / (…, tol, min, max, pow), blockTensor/LinearAlgebra.h:87,115,142"""
    hu, hd, hv = vp(), vp(), vp()
    trunc = tol is not None
    _check(a.ctx.lib.qtb_svd(a.ctx.h, a.h, int(split), 1 if trunc else 0, float(tol or 0.0), int(min_size),
                             -1 if max_size is None else int(max_size), float(pow), C.byref(hu), C.byref(hd),
                             C.byref(hv)))
    return BTensor(a.ctx, hu), BTensor(a.ctx, hd), BTensor(a.ctx, hv)


def eigh(a: BTensor, split: int, tol: Optional[float] = None, min_size: int = 1, max_size: Optional[int] = None,
         pow: float = 1.0):
    """reference quantit::eigh(btensor, split[, tol, min, max, pow]), blockTensor/LinearAlgebra.h:167-192: (e, U) with
    A = U diag(e) U^T per charge group, e ascending inside a group"""
    he, hu = vp(), vp()
    trunc = tol is not None
    _check(a.ctx.lib.qtb_eigh(a.ctx.h, a.h, int(split), 1 if trunc else 0, float(tol or 0.0), int(min_size),
                              -1 if max_size is None else int(max_size), float(pow), C.byref(he), C.byref(hu)))
    return BTensor(a.ctx, he), BTensor(a.ctx, hu)


def truncate(U: Optional[BTensor], d: BTensor, V: Optional[BTensor], max_size: Optional[int], min_size: int, tol: float,
             pow: float = 2.0):
    """reference quantit::truncate(U, d, V, max, min, tol, pow), blockTensor/LinearAlgebra.h:244-247"""
    hu, hd, hv = vp(), vp(), vp()
    _check(d.ctx.lib.qtb_truncate(d.ctx.h, U.h if U is not None else None, d.h, V.h if V is not None else None,
                                  -1 if max_size is None else int(max_size), int(min_size), float(tol), float(pow),
                                  C.byref(hu), C.byref(hd), C.byref(hv)))
    return (BTensor(d.ctx, hu) if U is not None else None, BTensor(d.ctx, hd), BTensor(d.ctx, hv) if V is not None else None)


def hamil2site_times_state(state: BTensor, hamil: BTensor, lenv: BTensor, renv: BTensor) -> BTensor:
    """reference details::hamil2site_times_state, dmrg.h:61, dmrg.cpp:520"""
    h = vp()
    _check(state.ctx.lib.qtb_heff_apply(state.ctx.h, state.h, hamil.h, lenv.h, renv.h, C.byref(h)))
    return BTensor(state.ctx, h)


def compute_left_env(hamil: BTensor, mps: BTensor, left_env: BTensor) -> BTensor:
    """reference compute_left_env, dmrg.cpp:424-459"""
    h = vp()
    _check(mps.ctx.lib.qtb_env_left(mps.ctx.h, hamil.h, mps.h, left_env.h, C.byref(h)))
    return BTensor(mps.ctx, h)


def compute_right_env(hamil: BTensor, mps: BTensor, right_env: BTensor) -> BTensor:
    """reference compute_right_env, dmrg.cpp:468-493"""
    h = vp()
    _check(mps.ctx.lib.qtb_env_right(mps.ctx.h, hamil.h, mps.h, right_env.h, C.byref(h)))
    return BTensor(mps.ctx, h)


def two_sites_update(state: BTensor, hamil: BTensor, lenv: BTensor, renv: BTensor):
    """reference two_sites_update, dmrg.cpp:623-651: returns (energy, updated state)"""
    h = vp()
    e = C.c_double()
    _check(state.ctx.lib.qtb_two_sites_update(state.ctx.h, state.h, hamil.h, lenv.h, renv.h, C.byref(e), C.byref(h)))
    return e.value, BTensor(state.ctx, h)


class dmrg_options:
    """reference include/dmrg_options.h:15-58 (same field names and defaults)"""

    def __init__(self, cutoff: float = 1e-6, convergence_criterion: float = 1e-5, maximum_bond: Optional[int] = None,
                 minimum_bond: int = 4, maximum_iterations: int = 1000):
        self.cutoff = cutoff
        self.convergence_criterion = convergence_criterion
        self.maximum_bond = maximum_bond
        self.minimum_bond = minimum_bond
        self.maximum_iterations = maximum_iterations


def dmrg(hamiltonian: List[BTensor], in_out_state: List[BTensor], options: dmrg_options, oc: int = 0, log: Optional[dict] = None):
    """reference quantit::dmrg(bMPO&, bMPS&, const dmrg_options&, dmrg_logger&), dmrg.h:39: optimises `in_out_state`
    in place (the list's BTensor handles are updated) and returns the energy. `log`, if given, receives per-sweep
    energies / seconds / mid-chain bond dimension (what dmrg_log_sweeptime records)."""
    ctx = in_out_state[0].ctx
    L = len(hamiltonian)
    H = (vp * L)(*[t.h for t in hamiltonian])
    P = (vp * L)(*[t.h for t in in_out_state])
    o = DmrgOptionsC(options.cutoff, options.convergence_criterion,
                     -1 if options.maximum_bond is None else int(options.maximum_bond), int(options.minimum_bond),
                     int(options.maximum_iterations))
    n = int(options.maximum_iterations)
    se, ss, sb = (C.c_double * n)(), (C.c_double * n)(), (C.c_int64 * n)()
    E, ns, occ = C.c_double(), i64(), i64(oc)
    _check(ctx.lib.qtb_dmrg(ctx.h, L, H, P, C.byref(occ), C.byref(o), C.byref(E), C.byref(ns), se, ss, sb))
    if log is not None:
        log["energy"] = [se[i] for i in range(ns.value)]
        log["seconds"] = [ss[i] for i in range(ns.value)]
        log["mid_bond"] = [sb[i] for i in range(ns.value)]
        log["oc"] = occ.value
    return E.value


DMRG_LOG_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.POINTER(C.c_int64), C.c_int64)


def dmrg_logged(hamiltonian: List[BTensor], in_out_state: List[BTensor], options: dmrg_options, logger, oc: int = 0):
    """quantit::dmrg with a dmrg_logger (reference include/dmrg_logger.h): `logger(iteration, energy, seconds, bond_dims)`
    is called after every sweep while the run is in progress. Returns (energy, sweeps, final centre)."""
    ctx = in_out_state[0].ctx
    L = len(hamiltonian)
    H = (vp * L)(*[t.h for t in hamiltonian])
    P = (vp * L)(*[t.h for t in in_out_state])
    o = DmrgOptionsC(options.cutoff, options.convergence_criterion,
                     -1 if options.maximum_bond is None else int(options.maximum_bond), int(options.minimum_bond),
                     int(options.maximum_iterations))

    def _cb(user, it, e, secs, bonds, n):
        try:
            logger(int(it), float(e), float(secs), [int(bonds[i]) for i in range(n)])
        except Exception:  # no exception may cross the C ABI
            import traceback
            traceback.print_exc()

    cb = DMRG_LOG_FN(_cb)
    e, ns, occ = C.c_double(), i64(), i64(oc)
    _check(ctx.lib.qtb_dmrg_logged(ctx.h, L, H, P, C.byref(occ), C.byref(o), C.byref(e), C.byref(ns),
                                   C.cast(cb, C.c_void_p), None))
    return e.value, int(ns.value), int(occ.value)


def contract(a: List[BTensor], b: List[BTensor], obs: Optional[List[BTensor]] = None) -> float:
    """reference quantit::contract(const bMPS&, const bMPS&[, const bMPO&]), MPT.h:724-727: <a|obs|b> or <a|b> (b enters
    conjugated, identity / all-ones edges on the outer bonds). Returns the scalar."""
    ctx = a[0].ctx
    L = len(a)
    if len(b) != L or (obs is not None and len(obs) != L):
        raise ValueError("contract: the chains must have the same length")
    A = (vp * L)(*[t.h for t in a])
    B = (vp * L)(*[t.h for t in b])
    H = (vp * L)(*[t.h for t in obs]) if obs is not None else None
    out = C.c_double()
    _check(ctx.lib.qtb_contract(ctx.h, L, A, B, H, C.byref(out)))
    return out.value


def move_oc(state: List[BTensor], oc: int, target: int) -> int:
    """reference bMPS::move_oc(int), MPT.h:611: gauge walk of the orthogonality centre from `oc` to `target` with
    untruncated block SVDs; the BTensor handles in `state` are updated in place. Returns the new centre."""
    ctx = state[0].ctx
    L = len(state)
    P = (vp * L)(*[t.h for t in state])
    occ = i64(oc)
    _check(ctx.lib.qtb_move_oc(ctx.h, L, P, C.byref(occ), int(target)))
    return int(occ.value)


def coalesce(mpo: List[BTensor], cutoff: float) -> List[BTensor]:
    """reference bMPO::coalesce(cutoff), MPT.h:699: compresses the MPO bonds with truncated block SVDs, in place (the
    BTensor handles in `mpo` are updated); returns the list."""
    ctx = mpo[0].ctx
    L = len(mpo)
    P = (vp * L)(*[t.h for t in mpo])
    _check(ctx.lib.qtb_coalesce(ctx.h, L, P, float(cutoff)))
    return mpo


def tensordot_host(a: dict, b: dict, dims_a, dims_b, ctx: Optional[Context] = None):
    """End-to-end C-ABI call from host buffers to host buffers (qtb_tensordot_host): `a`/`b` are dicts with keys
    sec_sizes, cvals, sel, blocks (and optionally mods, or prebuilt flat arrays under '_flat'). Returns
    (block indices, flat fp64 data)."""
    ctx = ctx or default_context()
    fa, fb = flatten_host(a), flatten_host(b)
    da, db = _arr(dims_a), _arr(dims_b)
    nob, numel = i64(), i64()
    args = [ctx.h, fa["nc"], _pi(fa["mods"]) if fa["mods"] is not None else None,
            fa["rank"], _pi(fa["nsec"]), _pi(fa["ss"]), _pi(fa["cv"]), _pi(fa["sel"]), fa["nb"], _pi(fa["idx"]), _pf(fa["data"]),
            fb["rank"], _pi(fb["nsec"]), _pi(fb["ss"]), _pi(fb["cv"]), _pi(fb["sel"]), fb["nb"], _pi(fb["idx"]), _pf(fb["data"]),
            len(da), _pi(da), _pi(db)]
    _check(ctx.lib.qtb_tensordot_host(*args, C.byref(nob), C.byref(numel), None, None))
    r_out = fa["rank"] + fb["rank"] - 2 * len(da)
    cidx = np.zeros(max(nob.value * r_out, 1), np.int64)
    cdata = np.zeros(max(numel.value, 1), np.float64)
    _check(ctx.lib.qtb_tensordot_host(*args, C.byref(nob), C.byref(numel), _pi(cidx), _pf(cdata)))
    return cidx[:nob.value * r_out].reshape(nob.value, r_out), cdata[:numel.value]


def flatten_host(t: dict) -> dict:
    """flat int64/fp64 arrays of a host block tensor description (done once, outside any timed region)"""
    if "_flat" in t:
        return t["_flat"]
    keys = list(t["blocks"].keys())
    f = {
        "rank": len(t["sec_sizes"]), "nc": len(t["sel"]),
        "mods": _arr(t["mods"]) if t.get("mods") is not None else None,
        "nsec": _arr([len(s) for s in t["sec_sizes"]]),
        "ss": _arr([x for s in t["sec_sizes"] for x in s]),
        "cv": _arr([c for cs in t["cvals"] for q in cs for c in q]),
        "sel": _arr(t["sel"]), "nb": len(keys), "idx": _arr([i for k in keys for i in k]),
        "data": np.ascontiguousarray(np.concatenate([np.asarray(t["blocks"][k], np.float64).reshape(-1) for k in keys])
                                     if keys else np.zeros(1)),
    }
    t["_flat"] = f
    return f
