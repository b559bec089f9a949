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


def flat_map_merge_plan(keys_a: List[Index], keys_b: List[Index]):
    """Literal restatement of flat_map::insert(first, last, collision, nocollision), reference
    include/blockTensor/flat_map.h:350-425, run on keys only. Returns, for every key of the merged (sorted) list,
    (key, pa, pb): pa = power of alpha applied to this's element (None if absent), pb = power of alpha applied to the
    other's element (None if absent). The ideal result is pa = 0, pb = 1 everywhere; the reference's final
    `for_each(begin(), move_to_end, nocollision)` also hits ORIGINAL elements whose keys are smaller than every key of
    the other map, so those get pa = 1 (multiplied by alpha) — observed behaviour, reproduced here."""
    import bisect
    content = [[k, {"a": 0}] for k in keys_a]           # value = {source: power of alpha}
    other = [[k, {"b": 0}] for k in keys_b]
    keyof = lambda lst: [e[0] for e in lst]
    first, last = 0, len(other)
    n1 = last - first
    look = 0
    for i in range(first, last):
        look = bisect.bisect_left(keyof(content), other[i][0], look, len(content))
        if look == len(content):
            break
        if not (other[i][0] < content[look][0]):
            n1 -= 1
    content.extend([[None, {}] for _ in range(n1)])
    look = len(content) - n1
    move_to_end = len(content)

    def nocollision(e):
        for src in e[1]:
            e[1][src] += 1

    def collision(x, y):   # a.add_(b, alpha)
        for src, pw in y[1].items():
            x[1][src] = pw + 1

    while look != 0 and last != first:
        ck = keyof(content[:look])
        la = bisect.bisect_left(keyof(other), content[look - 1][0], first, last)
        fc = 1 if (la != last and not (content[look - 1][0] < other[la][0])) else 0
        cnt = last - (la + fc)
        for t in range(cnt):
            src = other[last - 1 - t]
            content[move_to_end - 1 - t] = [src[0], dict(src[1])]
        old = move_to_end
        move_to_end -= cnt
        for e in content[move_to_end:old]:
            nocollision(e)
        if fc:
            move_to_end -= 1
            look -= 1
            if move_to_end != look:
                content[move_to_end] = content[look]
            collision(content[move_to_end], other[la])
        last = la
        if last == first:
            break
        lb = bisect.bisect_left(keyof(content[:look]), other[last - 1][0], 0, look)
        if move_to_end != look:
            seg = content[lb:look]
            content[move_to_end - len(seg):move_to_end] = seg
        move_to_end -= look - lb
        look = lb
        fc = 1 if (last != first and not (other[last - 1][0] < content[move_to_end][0])) else 0
        if fc:
            collision(content[move_to_end], other[last - 1])
        last -= fc
    rem = other[first:last]
    if rem:
        assert move_to_end == len(rem)
        content[move_to_end - len(rem):move_to_end] = [[e[0], dict(e[1])] for e in rem]
    for e in content[0:move_to_end]:
        nocollision(e)
    return [(tuple(k), v.get("a"), v.get("b")) for k, v in content]


def add(a: BT, b: BT, alpha: float = 1.0) -> BT:
    """reference btensor::add / add_, sources/btensor.cpp:2666-2752: flat_map::merge of the block lists, a + alpha*b
    on collisions, blocks present on one side only are copied (those of b scaled by alpha). The merge is restated
    literally (flat_map_merge_plan) because the reference's merge also scales by alpha those blocks of `a` that sort
    before every block of `b`."""
    assert a.sec_sizes == b.sec_sizes and a.cvals == b.cvals and a.sel == b.sel, "add: structure mismatch"
    out = a.structure_like()
    for key, pa, pb in flat_map_merge_plan(sorted(a.blocks), sorted(b.blocks)):
        v = 0.0
        if pa is not None:
            v = v + (alpha ** pa) * a.blocks[key]
        if pb is not None:
            v = v + (alpha ** pb) * b.blocks[key]
        out.blocks[key] = np.array(v, dtype=np.float64, copy=True)
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


# ----------------------------------------------------------------------------------------------------------------
# block SVD and truncation
# ----------------------------------------------------------------------------------------------------------------
def reshape_split(t: BT, split: int) -> BT:
    """reference btensor::reshape({split}), sources/btensor.cpp:2986-3024 with reshape_helpers :2760-2857:
    rank-2 tensor whose row (col) sections enumerate ALL combinations of the sections of dims [0,split)
    ([split,rank)) in row-major order (last index fastest); size = product, charge = product."""
    groups = [list(range(0, split)), list(range(split, t.rank))]
    sec_sizes, cvals = [], []
    for g in groups:
        sizes, cv = [], []
        for combo in (np.ndindex(*[len(t.sec_sizes[d]) for d in g]) if g else [()]):
            s, q = 1, q_neutral(t.nc)
            for d, i in zip(g, combo):
                s *= t.sec_sizes[d][i]
                q = q_op(q, t.cvals[d][i], t.mods)
            sizes.append(s)
            cv.append(q)
        sec_sizes.append(sizes)
        cvals.append(cv)
    out = BT(sec_sizes, cvals, t.sel, {}, t.mods)
    for idx, blk in t.blocks.items():
        new = []
        for g in groups:
            f = 0
            for d in g:
                f = f * len(t.sec_sizes[d]) + idx[d]
            new.append(f)
        rows = int(np.prod([blk.shape[d] for d in groups[0]])) if groups[0] else 1
        out.blocks[tuple(new)] = blk.reshape(rows, -1)
    return out


def reorder_by_cvals(t: BT) -> List[Index]:
    """reference LA_helpers::reorder_by_cvals, sources/btensor_linalg.cpp:30-67: stable sort of the (lexicographically
    ordered) block list by (row-section charge, col-section charge) under the tuple-lexicographic '<'."""
    r = t.rank
    items = [k for k, _ in t.sorted_items()]
    return sorted(items, key=lambda k: (t.cvals[r - 2][k[r - 2]], t.cvals[r - 1][k[r - 1]]))  # sorted() is stable


def svd_groups(t: BT):
    """reference LA_helpers::compact_dense + compact_dense_single, btensor_linalg.cpp:82-255, for a rank-2 tensor:
    list of (dense matrix, rows [(section, slice)], cols [(section, slice)] sorted by section)."""
    order = reorder_by_cvals(t)
    groups: List[List[Index]] = []
    for k in order:
        key = (t.cvals[0][k[0]], t.cvals[1][k[1]])
        if groups and (t.cvals[0][groups[-1][-1][0]], t.cvals[1][groups[-1][-1][1]]) == key:
            groups[-1].append(k)
        else:
            groups.append([k])
    out = []
    for g in groups:
        rows, cols = [], {}
        racc = cacc = 0
        cur_row = None
        for k in g:
            if k[1] not in cols:
                n = t.blocks[k].shape[1]
                cols[k[1]] = slice(cacc, cacc + n)
                cacc += n
            if cur_row != k[0]:
                n = t.blocks[k].shape[0]
                rows.append((k[0], slice(racc, racc + n)))
                racc += n
                cur_row = k[0]
        dense = np.zeros((racc, cacc))
        rmap = dict(rows)
        for k in g:
            dense[rmap[k[0]], cols[k[1]]] = t.blocks[k]
        out.append((dense, rows, sorted(cols.items())))
    return out


def svd_rank2(t: BT):
    """reference svd(const btensor&, some=true, compute_uv=true), btensor_linalg.cpp:390-502."""
    assert t.rank == 2
    groups = svd_groups(t)
    nc, mods = t.nc, t.mods
    d_sizes = [min(g[0].shape) for g in groups]
    right_q = [t.cvals[1][g[2][0][0]] for g in groups]
    neutral = q_neutral(nc)
    d = BT([d_sizes], [[neutral] * len(groups)], neutral, {}, mods)
    U = BT([list(t.sec_sizes[0]), list(d_sizes)], [list(t.cvals[0]), list(right_q)], t.sel, {}, mods)
    V = BT([list(t.sec_sizes[1]), list(d_sizes)], [[q_inv(q, mods) for q in t.cvals[1]], list(right_q)], neutral, {},
           mods)
    for b_i, (dense, rows, cols) in enumerate(groups):
        u, s, vt = np.linalg.svd(dense, full_matrices=False)
        v = vt.T
        for sec, sl in rows:
            U.blocks[(sec, b_i)] = u[sl, :].copy()
        for sec, sl in cols:
            V.blocks[(sec, b_i)] = v[sl, :].copy()
        d.blocks[(b_i,)] = s.copy()
    return U, d, V


def _unflatten(f: int, nsecs: Sequence[int]) -> Tuple[int, ...]:
    out = []
    for n in reversed(nsecs):
        out.append(f % n)
        f //= n
    return tuple(reversed(out))


def svd(t: BT, split: int):
    """reference svd(const btensor&, size_t split), btensor_linalg.cpp:503-534: reshape -> rank-2 svd -> reshape_as."""
    rU, d, rV = svd_rank2(reshape_split(t, split))
    left, right = list(range(split)), list(range(split, t.rank))
    U = BT([list(t.sec_sizes[i]) for i in left] + [list(rU.sec_sizes[1])],
           [list(t.cvals[i]) for i in left] + [list(rU.cvals[1])], rU.sel, {}, t.mods)
    V = BT([list(t.sec_sizes[i]) for i in right] + [list(rV.sec_sizes[1])],
           [[q_inv(q, t.mods) for q in t.cvals[i]] for i in right] + [list(rV.cvals[1])], rV.sel, {}, t.mods)
    for (rs, b), blk in rU.blocks.items():
        idx = _unflatten(rs, [len(t.sec_sizes[i]) for i in left]) + (b,)
        U.blocks[idx] = blk.reshape(tuple(t.sec_sizes[i][j] for i, j in zip(left, idx[:-1])) + (blk.shape[1],))
    for (cs, b), blk in rV.blocks.items():
        idx = _unflatten(cs, [len(t.sec_sizes[i]) for i in right]) + (b,)
        V.blocks[idx] = blk.reshape(tuple(t.sec_sizes[i][j] for i, j in zip(right, idx[:-1])) + (blk.shape[1],))
    return U, d, V


def compute_last_index(vd: np.ndarray, tol: float, pw: float, min_size: int, max_size: int) -> int:
    """reference compute_last_index, sources/LinearAlgebra.cpp:57-75 (vd sorted descending)."""
    n = len(vd)
    toln = tol ** pw
    last = n - 1
    trunc = abs(vd[last]) ** pw
    while last >= min_size:
        if trunc > toln and last < max_size:
            break
        last -= 1
        trunc += abs(vd[last]) ** pw
    return last


def _remove_unit_blocks(keys: List[Index], sector: int) -> List[Index]:
    """literal restatement of the remove_unit_blocks lambda, btensor_linalg.cpp:688-706 (a compaction loop whose read
    cursor advances by at most one per step: of every run of consecutive removable blocks only every other one is
    dropped — observed behaviour, SURVEY.md appendix B spirit: parity is against what the reference does)."""
    lst = list(keys)
    src = dest = 0
    n = len(lst)
    while dest != n:
        dest += 1 if lst[dest][-1] == sector else 0
        if dest != src and dest != n:
            lst[dest], lst[src] = lst[src], lst[dest]
        dest += 1 if dest != n else 0
        src += 1
    return lst[: n - (dest - src)]


def truncate(U: BT, d: BT, V: BT, max_size: int, min_size: int, tol: float, pw: float = 2.0):
    """reference truncate_impl, btensor_linalg.cpp:657-755."""
    U, d, V = U.copy(), d.copy(), V.copy()
    items = d.sorted_items()
    vd = np.sort(np.concatenate([v for _, v in items]))[::-1] if items else np.zeros(0)
    last = compute_last_index(vd, tol, pw, min_size, max_size)
    thr = vd[last]
    thr -= 2 * thr * np.finfo(np.float64).eps
    for (sec,), db in reversed(items):
        n = 0
        while n < len(db) and db[n] > thr:  # lower_bound_impl2, btensor_linalg.cpp:548-558
            n += 1
        if n == 0:
            for T in (U, V):
                keep = _remove_unit_blocks([k for k, _ in T.sorted_items()], sec)
                T.blocks = {k: T.blocks[k] for k in keep}
            del d.blocks[(sec,)]
        else:
            d.blocks[(sec,)] = db[:n].copy()
            d.sec_sizes[0][sec] = n
            for T in (U, V):
                T.sec_sizes[-1][sec] = n
                for k in list(T.blocks):
                    if k[-1] == sec:
                        T.blocks[k] = T.blocks[k][..., :n].copy()
    return U, d, V


def svd_trunc(t: BT, split: int, tol: float, min_size: int, max_size: int, pw: float = 2.0):
    """reference svd(A, split, tol, min, max, pow) = truncate(svd(A, split), max, min, tol, pow), btensor_linalg.cpp:805."""
    U, d, V = svd(t, split)
    return truncate(U, d, V, max_size, min_size, tol, pw)


def reshape(t: BT, index_groups: Sequence[int]) -> BT:
    """reference btensor::reshape(index_groups), sources/btensor.cpp:2986-3024 (helpers :2832-2985): consecutive dims
    between the boundaries are merged; every combination of the merged sections becomes a section (row-major, last
    index fastest) with size = product and charge = product; a block's new index is the flattened old one."""
    r = t.rank
    bounds = [0] + [int(x) for x in index_groups] + [r]
    out_rank = len(bounds) - 1
    sec_sizes, cvals = [], []
    for g in range(out_rank):
        dims = list(range(bounds[g], bounds[g + 1]))
        ss, cv = [1], [q_neutral(t.nc)]
        for d in dims:
            ss = [a * b for a in ss for b in t.sec_sizes[d]]
            cv = [q_op(a, b, t.mods) for a in cv for b in t.cvals[d]]
        sec_sizes.append(ss)
        cvals.append(cv)
    out = BT(sec_sizes, cvals, t.sel, {}, t.mods)
    for idx, blk in t.blocks.items():
        new = []
        for g in range(out_rank):
            f = 0
            for d in range(bounds[g], bounds[g + 1]):
                f = f * t.nsec[d] + idx[d]
            new.append(f)
        shape = [int(np.prod(blk.shape[bounds[g]:bounds[g + 1]], dtype=np.int64)) for g in range(out_rank)]
        out.blocks[tuple(new)] = np.ascontiguousarray(blk).reshape(shape)
    return out


def tensorgdot(c: BT, a: BT, b: BT, dims_a: Sequence[int], dims_b: Sequence[int], beta: float = 1.0, alpha: float = 1.0) -> BT:
    """D = alpha*C + beta*A.B: btensor::tensorgdot is declared (reference include/blockTensor/btensor.h:624-627) and never
    defined; semantics of the dense routine, include/tensorgdot.h:22-86 (argument order add, mul1, mul2, dims1, dims2,
    beta, alpha). Block list = union of C's and (A.B)'s."""
    t = tensordot(a, b, dims_a, dims_b)
    out = c.structure_like()
    for k in sorted(set(c.blocks) | set(t.blocks)):
        v = 0.0
        if k in c.blocks:
            v = v + alpha * c.blocks[k]
        if k in t.blocks:
            v = v + beta * t.blocks[k]
        out.blocks[k] = v
    return out


def eigh_groups(t: BT, split: int):
    """Mathematical oracle for eigh(btensor, split) (reference blockTensor/LinearAlgebra.h:159-192, btensor_linalg.cpp:
    294-389 — the reference's own implementation crashes in the oracle build, see ref_harness eigh): numpy.linalg.eigh of
    every densified charge group of the rank-2 reshape, groups in the order of svd_groups. Returns [(e ascending, U)]"""
    m = reshape_split(t, split)
    out = []
    for dense, rows, cols in svd_groups(m):
        e, u = np.linalg.eigh((dense + dense.T) / 2)
        out.append((e, u, rows, cols))
    return out


def recompose(U: BT, d: BT, V: BT) -> BT:
    """U * d * V^T contracted over the bond: the gauge-independent quantity SVD parity is judged on."""
    Ud = mul_bcast(U, d)
    Vc = conj(V)
    return tensordot(Ud, Vc, [U.rank - 1], [V.rank - 1])


# ----------------------------------------------------------------------------------------------------------------
# two-site DMRG pieces (reference sources/dmrg.cpp)
# ----------------------------------------------------------------------------------------------------------------
def hamil2site_times_state(state: BT, hamil: BT, lenv: BT, renv: BT) -> BT:
    """dmrg.cpp:520-531"""
    out = tensordot(lenv, state, [0], [0])
    out = tensordot(out, hamil, [0, 2, 3], [0, 4, 5])
    return tensordot(out, renv, [1, 4], [0, 1])


def compute_left_env(hamil: BT, mps: BT, left_env: BT) -> BT:
    """dmrg.cpp:424-459"""
    out = tensordot(left_env, mps, [0], [0])
    out = tensordot(out, hamil, [0, 2], [0, 3])
    return tensordot(out, conj(mps), [0, 2], [0, 1])


def compute_right_env(hamil: BT, mps: BT, right_env: BT) -> BT:
    """dmrg.cpp:468-493"""
    out = tensordot(right_env, mps, [0], [2])
    out = tensordot(out, hamil, [0, 3], [2, 3])
    return tensordot(out, conj(mps), [3, 0], [1, 2])


def compute_2sites_hamil(w1: BT, w2: BT) -> BT:
    """dmrg.cpp:503-515"""
    return permute(tensordot(w1, w2, [2], [0]), [0, 1, 3, 4, 2, 5])


def eig2x2(a0: float, a1: float, b: float):
    """eig2x2Mat_impl, dmrg.cpp:543-572"""
    crit = np.sqrt((a0 - a1) ** 2 + 4 * (b * b))
    E0 = (a0 + a1 - crit) / 2
    delt = E0 - a1
    with np.errstate(all="ignore"):
        o = np.sqrt(np.float64(delt) / np.float64(-crit))
        zero_o = bool((o + E0) == E0) or bool(np.isnan(o))
        n = (b * o) / np.float64(delt)
    if zero_o:
        n, o = 1.0, 0.0
    if np.isnan(o) or np.isnan(n):
        raise ArithmeticError("nan found in output tensor")
    return float(E0), float(o), float(n)


def two_sites_update(state: BT, hamil: BT, lenv: BT, renv: BT):
    """one_step_lanczos_impl + two_sites_update_impl, dmrg.cpp:584-638. Returns (E, updated state)."""
    psi_ip = hamil2site_times_state(state, hamil, lenv, renv)
    a0 = dot_all(psi_ip, conj(state))
    psi_ip = add(psi_ip, mul_bcast(state, scalar(a0, state.nc, state.mods)), -1.0)  # psi_ip -= state * a0
    b = float(np.sqrt(dot_all(psi_ip, conj(psi_ip))))
    if abs(b) >= 1e-15:
        for k in psi_ip.blocks:
            psi_ip.blocks[k] = psi_ip.blocks[k] / b
    a1 = dot_all(conj(psi_ip), hamil2site_times_state(psi_ip, hamil, lenv, renv))
    E, o, n = eig2x2(a0, a1, b)
    out = state.structure_like()
    for k, v in state.blocks.items():
        out.blocks[k] = o * v
    return E, add(out, psi_ip, n)


def dmrg_two_site_step(mps: List[BT], mpo: List[BT], h2: List[BT], env: Dict[int, BT], oc: int, step: int, cutoff: float,
                       min_bond: int, max_bond: int):
    """dmrg_2sites_update::operator(), dmrg.cpp:163-206. Mutates mps/env; returns (E, new oc)."""
    theta = tensordot(mps[oc], mps[oc + 1], [2], [0])
    E, theta = two_sites_update(theta, h2[oc], env[oc - 1], env[oc + 2])
    u, d, v = svd_trunc(theta, 2, cutoff, min_bond, max_bond)
    nrm = np.sqrt(sum(float(np.sum(x * x)) for x in d.blocks.values()))
    for k in d.blocks:
        d.blocks[k] = d.blocks[k] / nrm
    if step == 1:
        mps[oc] = u
        mps[oc + 1] = permute(conj(mul_bcast(v, d)), [2, 0, 1])
        env[oc] = compute_left_env(mpo[oc], mps[oc], env[oc - 1])
    else:
        mps[oc] = mul_bcast(u, d)
        mps[oc + 1] = permute(conj(v), [2, 0, 1])
        env[oc + 1] = compute_right_env(mpo[oc + 1], mps[oc + 1], env[oc + 2])
    return E, oc + step


def trivial_edges(mps: List[BT], mpo: List[BT]):
    """generate_env_impl's edge tensors, dmrg.cpp:370-392: ones on 1x1x1 legs (inverse ket charge, inverse MPO charge,
    ket charge)."""
    def edge(state_leg, ham_leg):
        (ss, sq), (hs, hq) = state_leg, ham_leg
        mods = mps[0].mods
        t = BT([list(ss), list(hs), list(ss)], [[q_inv(q, mods) for q in sq], [q_inv(q, mods) for q in hq], list(sq)],
               q_neutral(mps[0].nc), {}, mods)
        for idx in all_allowed_indices(t):
            t.blocks[idx] = np.ones(t.block_dims(idx))
        return t
    left = edge((mps[0].sec_sizes[0], mps[0].cvals[0]), (mpo[0].sec_sizes[0], mpo[0].cvals[0]))
    right = edge((mps[-1].sec_sizes[2], mps[-1].cvals[2]), (mpo[-1].sec_sizes[2], mpo[-1].cvals[2]))
    return left, right


def _eye_edge(a: BT, dim_a: int, b: BT, dim_b: int, h: BT | None, dim_h: int) -> BT:
    """eye_like(shape_from(A[dim]^-1, B[dim])) (reference btensor.cpp:2444-2458: identity blocks on the allowed (i, i)
    section pairs), times ones_like(edge_shape_prep(H, dim)) permuted {0,2,1} when an MPO is present
    (reference MPT.cpp:218-225)."""
    mods = a.mods
    if h is None:
        t = BT([list(a.sec_sizes[dim_a]), list(b.sec_sizes[dim_b])],
               [[q_inv(q, mods) for q in a.cvals[dim_a]], list(b.cvals[dim_b])], q_neutral(a.nc), {}, mods)
        for i in range(min(t.nsec)):
            if t.allowed((i, i)):
                t.blocks[(i, i)] = np.eye(*t.block_dims((i, i)))
        return t
    t = BT([list(a.sec_sizes[dim_a]), list(h.sec_sizes[dim_h]), list(b.sec_sizes[dim_b])],
           [[q_inv(q, mods) for q in a.cvals[dim_a]], [q_inv(q, mods) for q in h.cvals[dim_h]], list(b.cvals[dim_b])],
           q_neutral(a.nc), {}, mods)
    for i in range(min(t.nsec[0], t.nsec[2])):
        for j in range(t.nsec[1]):
            if t.allowed((i, j, i)):
                da, dw, db = t.block_dims((i, j, i))
                t.blocks[(i, j, i)] = np.einsum("ab,w->awb", np.eye(da, db), np.ones(dw))
    return t


def contract(a: List[BT], b: List[BT], obs: List[BT] | None = None) -> float:
    """reference contract(const bMPS&, const bMPS&, const bMPO&), sources/MPT.cpp:211-233, and
    contract(const bMPS&, const bMPS&), :275-292: three (two) tensordots per site, b conjugated, closed with the right
    edge; returns the value of the rank-0 result."""
    assert len(a) == len(b)
    left = _eye_edge(a[0], 0, b[0], 0, obs[0] if obs else None, 0)
    right = _eye_edge(a[-1], 2, b[-1], 2, obs[-1] if obs else None, 2)
    for i in range(len(a)):
        left = tensordot(left, a[i], [0], [0])
        if obs:
            left = tensordot(left, obs[i], [0, 2], [0, 3])
            left = tensordot(left, conj(b[i]), [0, 2], [0, 1])
        else:
            left = tensordot(left, conj(b[i]), [0, 1], [0, 1])
    res = tensordot(left, right, [0, 1, 2], [0, 1, 2]) if obs else tensordot(left, right, [0, 1], [0, 1])
    return res.item()


def move_oc(mps: List[BT], oc: int, target: int) -> int:
    """reference bMPS::move_oc(int), sources/MPT.cpp:75-111 (in place on the list; returns the new centre)"""
    if not (0 <= target < len(mps)):
        raise ValueError(" Proposed orthogonality center falls outside the MPS")
    while target < oc:
        u, d, v = svd(mps[oc], 1)
        mps[oc] = permute(conj(v), [2, 0, 1])
        mps[oc - 1] = tensordot(mps[oc - 1], mul_bcast(u, d), [2], [0])
        oc -= 1
    while target > oc:
        u, d, v = svd(mps[oc], 2)
        mps[oc] = u
        mps[oc + 1] = tensordot(conj(mul_bcast(v, d)), mps[oc + 1], [0], [0])
        oc += 1
    return oc


def coalesce(mpo: List[BT], cutoff: float) -> List[BT]:
    """reference bMPO::coalesce(cutoff), sources/MPT.cpp:154-168 (in place on the list): svd(tens, 3, cutoff) is the
    (tol, pow = 2) overload with min_size 1 and no maximum (btensor_linalg.cpp:811-816)"""
    for i in range(len(mpo) - 1):
        tens = permute(mpo[i], [0, 1, 3, 2])
        U, d, V = svd_trunc(tens, 3, cutoff, 1, 2 ** 62, 2.0)
        mpo[i + 1] = tensordot(conj(V), mpo[i + 1], [0], [0])
        mpo[i] = permute(mul_bcast(U, d), [0, 1, 3, 2])
    return mpo


def dmrg(mps: List[BT], mpo: List[BT], oc: int, cutoff: float, conv: float, max_bond: int, min_bond: int = 4,
         max_iter: int = 1000, log=None):
    """details::dmrg_impl + generate_env + sweep, dmrg.cpp:92-100,127-142,219-273,370-409. Returns the energy."""
    L = len(mpo)
    env: Dict[int, BT] = {}
    env[-1], env[L] = trivial_edges(mps, mpo)
    for i in range(oc):
        env[i] = compute_left_env(mpo[i], mps[i], env[i - 1])
    for i in range(L - 1, oc, -1):
        env[i] = compute_right_env(mpo[i], mps[i], env[i + 1])
    h2 = [compute_2sites_hamil(mpo[i], mpo[i + 1]) for i in range(L - 1)]
    E0 = 100000.0
    n_step = len(h2) - 1 + (1 if len(h2) == 1 else 0)
    step = 1 if oc == 0 else -1
    if len(h2) == 1:
        step = 0
    if oc == L - 1:
        oc -= 1
    for it in range(max_iter):
        E = None
        for _ in range(2 * n_step):
            E, oc = dmrg_two_site_step(mps, mpo, h2, env, oc, step, cutoff, min_bond, max_bond)
            if oc == 0 or oc == L - 2:
                step = -step
        if log:
            log(it, E, mps)
        E0, Et = E, E0
        if not (abs((E0 - Et) / E0) > conv):
            break
    return E0
