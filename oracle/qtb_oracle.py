"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of the reference's block-sparse hot path.

This module is the ORACLE: a plain restatement of what AlexandreFoley/QuantiT computes on the path named by
BASELINE.json:north_star (btensor::tensordot, block svd + truncation, two-site H_eff·psi, one-step Lanczos,
environment updates). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.
The product (quantit_b200 + libqtb.so) never does.

Parity status: PINNED. Every function below is checked in tests/test_oracle_vs_reference.py against
 (a) the golden vectors of the reference's own in-header tests (SVD grouping order,
     include/blockTensor/LinearAlgebra.h:276-277,309-310), and
 (b) outputs of the reference itself (oracle/_ref/ref_harness, the unmodified reference sources compiled by
     oracle/Makefile) committed as fixtures under tests/golden/ with the generating script
     tests/golden/make_golden.py.

Each function cites the reference file:line it restates. Charges are tuples of ints; a charge *type* is the tuple of
moduli (0 = Z, N>0 = C<N>), reference include/Conserved/quantity.h:62-239.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

Charge = Tuple[int, ...]
Index = Tuple[int, ...]


# ----------------------------------------------------------------------------------------------------------------
# charges  (reference include/Conserved/quantity.h:128-143,200-215 ; Composite/quantity_impl.h:326-329 for '<')
# ----------------------------------------------------------------------------------------------------------------
def q_op(a: Charge, b: Charge, mods: Sequence[int] | None = None) -> Charge:
    if mods is None:
        return tuple(x + y for x, y in zip(a, b))
    return tuple((x + y) % m if m else x + y for x, y, m in zip(a, b, mods))


def q_inv(a: Charge, mods: Sequence[int] | None = None) -> Charge:
    if mods is None:
        return tuple(-x for x in a)
    return tuple((-x) % m if m else -x for x, m in zip(a, mods))


def q_neutral(nc: int) -> Charge:
    return (0,) * nc


# ----------------------------------------------------------------------------------------------------------------
# block tensor  (reference include/blockTensor/btensor.h:105-110,821-837)
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class BT:
    """rank, sections per dim, section sizes per dim, section charges per dim, selection rule, sorted block map."""

    sec_sizes: List[List[int]]
    cvals: List[List[Charge]]
    sel: Charge
    blocks: Dict[Index, np.ndarray] = field(default_factory=dict)
    mods: Tuple[int, ...] | None = None

    @property
    def rank(self) -> int:
        return len(self.sec_sizes)

    @property
    def nsec(self) -> List[int]:
        return [len(s) for s in self.sec_sizes]

    @property
    def nc(self) -> int:
        return len(self.sel)

    def sizes(self) -> List[int]:
        return [sum(s) for s in self.sec_sizes]

    def block_dims(self, idx: Index) -> Tuple[int, ...]:
        return tuple(self.sec_sizes[d][i] for d, i in enumerate(idx))

    def allowed(self, idx: Index) -> bool:
        """reference btensor::block_conservation_rule_test, sources/btensor.cpp:333-345"""
        q = q_neutral(self.nc)
        for d, i in enumerate(idx):
            q = q_op(q, self.cvals[d][i], self.mods)
        return q == self.sel

    def sorted_items(self):
        return sorted(self.blocks.items(), key=lambda kv: kv[0])

    def structure_like(self) -> "BT":
        return BT([list(s) for s in self.sec_sizes], [list(c) for c in self.cvals], self.sel, {}, self.mods)

    def copy(self) -> "BT":
        out = self.structure_like()
        out.blocks = {k: v.copy() for k, v in self.blocks.items()}
        return out

    def numel_stored(self) -> int:
        return sum(int(v.size) for v in self.blocks.values())

    def to_dense(self) -> np.ndarray:
        """reference btensor::to_dense, sources/btensor.cpp:2491-2510"""
        out = np.zeros(self.sizes())
        offs = [np.concatenate([[0], np.cumsum(s)]) for s in self.sec_sizes]
        for idx, blk in self.blocks.items():
            sl = tuple(slice(offs[d][i], offs[d][i + 1]) for d, i in enumerate(idx))
            out[sl] = blk
        return out

    def item(self) -> float:
        assert self.rank == 0 or all(s == 1 for s in self.sizes())
        if not self.blocks:
            return 0.0
        return float(next(iter(self.blocks.values())).reshape(-1)[0])


def all_allowed_indices(t: BT) -> List[Index]:
    out = []
    for idx in np.ndindex(*t.nsec) if t.rank else [()]:
        if t.allowed(tuple(int(i) for i in idx)):
            out.append(tuple(int(i) for i in idx))
    return sorted(out)


def rand_like(shape: BT, rng: np.random.Generator) -> BT:
    """All selection-rule-allowed blocks, uniform [0,1) values (structure of reference rand_like,
    sources/btensor.cpp:2399-2404; the value stream is ours — inputs are exchanged as QTBT files)."""
    out = shape.structure_like()
    for idx in all_allowed_indices(shape):
        out.blocks[idx] = rng.random(shape.block_dims(idx))
    return out


def shape_from(*shapes: BT) -> BT:
    """tensor product of empty shapes: dims concatenated, selection rules multiplied
    (reference shape_from, sources/btensor.cpp:2267-2311)."""
    sec_sizes, cvals = [], []
    mods = shapes[0].mods
    sel = q_neutral(shapes[0].nc)
    for s in shapes:
        sec_sizes += [list(x) for x in s.sec_sizes]
        cvals += [list(x) for x in s.cvals]
        sel = q_op(sel, s.sel, mods)
    return BT(sec_sizes, cvals, sel, {}, mods)


# ----------------------------------------------------------------------------------------------------------------
# QTBT dump format (shared with oracle/ref_harness.cpp dump()/load())
# ----------------------------------------------------------------------------------------------------------------
def write_qtbt(t: BT, path: str) -> None:
    items = t.sorted_items()
    with open(path, "wb") as f:
        f.write(b"QTBT0001")
        f.write(struct.pack("<3q", t.rank, t.nc, len(items)))
        f.write(np.asarray(t.nsec, dtype="<i8").tobytes())
        for s in t.sec_sizes:
            f.write(np.asarray(s, dtype="<i8").tobytes())
        for c in t.cvals:
            f.write(np.asarray(c, dtype="<i8").reshape(-1).tobytes())
        f.write(np.asarray(t.sel, dtype="<i8").tobytes())
        for idx, _ in items:
            f.write(np.asarray(idx, dtype="<i8").tobytes())
        for _, blk in items:
            f.write(np.asarray(blk.shape, dtype="<i8").tobytes())
        for _, blk in items:
            f.write(np.ascontiguousarray(blk, dtype="<f8").tobytes())


def read_qtbt(path: str) -> BT:
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"QTBT0001", path
    pos = 8
    rank, nc, nb = struct.unpack_from("<3q", buf, pos)
    pos += 24

    def take(n):
        nonlocal pos
        a = np.frombuffer(buf, dtype="<i8", count=n, offset=pos)
        pos += 8 * n
        return a

    nsec = [int(x) for x in take(rank)]
    sec_sizes = [[int(x) for x in take(n)] for n in nsec]
    cvals = [[tuple(int(x) for x in take(nc)) for _ in range(n)] for n in nsec]
    sel = tuple(int(x) for x in take(nc))
    idxs = [tuple(int(x) for x in take(rank)) for _ in range(nb)]
    dims = [tuple(int(x) for x in take(rank)) for _ in range(nb)]
    out = BT(sec_sizes, cvals, sel, {})
    for idx, d in zip(idxs, dims):
        n = int(np.prod(d)) if len(d) else 1
        out.blocks[idx] = np.frombuffer(buf, dtype="<f8", count=n, offset=pos).reshape(d).copy()
        pos += 8 * n
    assert pos == len(buf), (pos, len(buf))
    return out


# ----------------------------------------------------------------------------------------------------------------
# structural ops
# ----------------------------------------------------------------------------------------------------------------
def permute(t: BT, perm: Sequence[int]) -> BT:
    """reference btensor::permute, sources/btensor.cpp:1754-1802 (blocks become views; here copies)."""
    perm = [p + t.rank if p < 0 else p for p in perm]
    out = BT([list(t.sec_sizes[p]) for p in perm], [list(t.cvals[p]) for p in perm], t.sel, {}, t.mods)
    for idx, blk in t.blocks.items():
        out.blocks[tuple(idx[p] for p in perm)] = np.transpose(blk, perm)
    return out


def conj(t: BT) -> BT:
    """reference btensor::conj = conj_only().inverse_cvals_(), sources/btensor.cpp:2156-2172. Real dtype: values
    unchanged, every section charge and the selection rule inverted."""
    out = BT([list(s) for s in t.sec_sizes], [[q_inv(c, t.mods) for c in cs] for cs in t.cvals], q_inv(t.sel, t.mods),
             dict(t.blocks), t.mods)
    return out


def inverse_cvals(t: BT) -> BT:
    return conj(t)


def tensordot(a: BT, b: BT, dims_a: Sequence[int], dims_b: Sequence[int]) -> BT:
    """reference btensor::tensordot, sources/btensor.cpp:1971-2119, with compute_tdot_shape :841-883,
    check_product_compat :783-825, compute_tdot_cval_sectSize :1908-1935, permute_bl :1843-1894.

    Output dims = free dims of a (ascending) then free dims of b (ascending); selection rule = sel(a)*sel(b);
    an output block exists iff at least one pair of blocks shares the contracted block indices; pairs are
    accumulated in ascending order of the contracted block index (first mm, then addmm_)."""
    dims_a = list(dims_a)
    dims_b = list(dims_b)
    if len(dims_a) != len(dims_b):
        raise ValueError("both dimension lists should have the same length.")
    if a.nc != b.nc or a.mods != b.mods:
        raise ValueError("the two tensors have different type of conserved quantities")
    for da, db in zip(dims_a, dims_b):
        if len(a.sec_sizes[da]) != len(b.sec_sizes[db]):
            raise ValueError("contracted dimensions need to match")
        for qa, qb in zip(a.cvals[da], b.cvals[db]):
            if q_op(qa, qb, a.mods) != q_neutral(a.nc):
                raise ValueError("contracted conserved numbers need to sum to zero")
    free_a = [i for i in range(a.rank) if i not in dims_a]
    free_b = [i for i in range(b.rank) if i not in dims_b]
    out = BT([list(a.sec_sizes[i]) for i in free_a] + [list(b.sec_sizes[i]) for i in free_b],
             [list(a.cvals[i]) for i in free_a] + [list(b.cvals[i]) for i in free_b],
             q_op(a.sel, b.sel, a.mods), {}, a.mods)
    k = len(dims_a)
    # permute_bl: key = [free..., contracted...] for a ; [free..., contracted...] for b (p2_prime), matrices
    # [prod(free), prod(contracted)] and [prod(contracted), prod(free)].
    ta = sorted(((tuple(i[d] for d in free_a), tuple(i[d] for d in dims_a),
                  np.transpose(blk, free_a + dims_a).reshape(int(np.prod([blk.shape[d] for d in free_a])), -1))
                 for i, blk in a.blocks.items()), key=lambda x: (x[0], x[1]))
    tb = sorted(((tuple(i[d] for d in free_b), tuple(i[d] for d in dims_b),
                  np.transpose(blk, dims_b + free_b).reshape(-1, int(np.prod([blk.shape[d] for d in free_b]))))
                 for i, blk in b.blocks.items()), key=lambda x: (x[0], x[1]))
    # group into "columns" (runs of equal free index); k == 0: every block is its own column (next_index special
    # case, btensor.cpp:2005-2009)
    def columns(lst):
        cols, cur = [], None
        for fr, ct, m in lst:
            if k == 0 or cur is None or cur[0] != fr:
                cur = (fr, [])
                cols.append(cur)
            cur[1].append((ct, m))
        return cols

    for fa, la in columns(ta):
        for fb, lb in columns(tb):
            # two-pointer merge on the contracted index (find_next_match, btensor.cpp:2013-2034)
            i = j = 0
            acc = None
            while i < len(la) and j < len(lb):
                if la[i][0] < lb[j][0]:
                    i += 1
                elif lb[j][0] < la[i][0]:
                    j += 1
                else:
                    if la[i][1].shape[1] != lb[j][1].shape[0]:
                        raise ValueError("mm shape mismatch (section sizes of contracted dims differ)")
                    prod = la[i][1] @ lb[j][1]
                    acc = prod if acc is None else acc + prod
                    i += 1  # "break the match": only the left iterator advances (btensor.cpp:2096,2103)
            if acc is not None:
                idx = fa + fb
                out.blocks[idx] = acc.reshape(out.block_dims(idx))
    return out


def tensordot_flops(a: BT, b: BT, dims_a: Sequence[int], dims_b: Sequence[int]) -> int:
    """Algorithmic flops = sum over matched block pairs of 2*m*n*k."""
    free_a = [i for i in range(a.rank) if i not in dims_a]
    free_b = [i for i in range(b.rank) if i not in dims_b]
    by_key: Dict[Index, List[Tuple[int, int]]] = {}
    for i, blk in b.blocks.items():
        key = tuple(i[d] for d in dims_b)
        n = int(np.prod([blk.shape[d] for d in free_b])) if free_b else 1
        kk = int(np.prod([blk.shape[d] for d in dims_b])) if dims_b else 1
        by_key.setdefault(key, []).append((n, kk))
    fl = 0
    for i, blk in a.blocks.items():
        key = tuple(i[d] for d in dims_a)
        m = int(np.prod([blk.shape[d] for d in free_a])) if free_a else 1
        for n, kk in by_key.get(key, []):
            fl += 2 * m * n * kk
    return fl


# ----------------------------------------------------------------------------------------------------------------
# elementwise ops used by the Lanczos step and the DMRG bookkeeping
# ----------------------------------------------------------------------------------------------------------------
def mul_bcast(big: BT, small: BT) -> BT:
    """reference btensor::mul / broadcast_operation_, sources/btensor.cpp:1204-1302 with mul_helpers::shape_compute
    :980-1084, restricted to the two cases the hot path uses (dmrg.cpp:188-199,595,603,635):
      * small.rank == 0 (a scalar btensor): every block of `big` is scaled; charges/selection rule multiplied by the
        scalar's (neutral on the path);
      * small.rank == 1 matched with big's LAST dim with identical sections: block (…,k) * block (k), blocks of `big`
        whose last index has no partner in `small` are dropped; section charges multiply."""
    out = big.structure_like()
    if small.rank == 0:
        out.sel = q_op(big.sel, small.sel, big.mods)
        if small.blocks:
            s = next(iter(small.blocks.values())).reshape(())
            for idx, blk in big.blocks.items():
                out.blocks[idx] = blk * s
        return out
    assert small.rank == 1 and small.sec_sizes[0] == big.sec_sizes[-1], "oracle mul_bcast: unsupported broadcast"
    out.cvals[-1] = [q_op(x, y, big.mods) for x, y in zip(big.cvals[-1], small.cvals[0])]
    out.sel = q_op(big.sel, small.sel, big.mods)
    for idx, blk in big.blocks.items():
        sb = small.blocks.get((idx[-1],))
        if sb is not None:
            out.blocks[idx] = blk * sb
    return out


def add(a: BT, b: BT, alpha: float = 1.0) -> BT:
    """reference btensor::add / add_, sources/btensor.cpp:2666-2752: sorted merge of the block lists, a + alpha*b on
    collisions, blocks present on one side only are copied (scaled by alpha when they come from b)."""
    assert a.sec_sizes == b.sec_sizes and a.cvals == b.cvals and a.sel == b.sel, "add: structure mismatch"
    out = a.structure_like()
    for idx, blk in a.blocks.items():
        out.blocks[idx] = blk.copy()
    for idx, blk in b.blocks.items():
        if idx in out.blocks:
            out.blocks[idx] = out.blocks[idx] + alpha * blk
        else:
            out.blocks[idx] = alpha * blk
    return out


def scalar(val: float, nc: int, mods=None) -> BT:
    return BT([], [], q_neutral(nc), {(): np.array(val, dtype=np.float64)}, mods)


def dot_all(a: BT, b: BT) -> float:
    """tensordot over every index pair i<->i: a rank-0 btensor (dmrg.cpp:593,596,605)."""
    r = tensordot(a, b, list(range(a.rank)), list(range(b.rank)))
    return r.item()


def same_structure(a: BT, b: BT) -> bool:
    return (a.sec_sizes == b.sec_sizes and a.cvals == b.cvals and a.sel == b.sel
            and sorted(a.blocks) == sorted(b.blocks)
            and all(a.blocks[k].shape == b.blocks[k].shape for k in a.blocks))


def max_rel_err(a: BT, b: BT) -> float:
    """max |a-b| over all elements / max |b| (norm-wise relative error, the fp64 parity measure)."""
    num, den = 0.0, 0.0
    for k in b.blocks:
        den = max(den, float(np.max(np.abs(b.blocks[k]))) if b.blocks[k].size else 0.0)
        num = max(num, float(np.max(np.abs(a.blocks[k] - b.blocks[k]))) if b.blocks[k].size else 0.0)
    return num / den if den else num
