"""sharded.py -- one multiply on the N GPUs of one box: one process per GPU, torch.distributed for
the plumbing (NCCL on GPUs; the same code runs over gloo in the CPU tests).

The reference planned exactly this split and never built it: "B is sliced, and the slices of C are
concatenated" (Source/GB_AxB_parallel.c:52), with the slices balanced by the cumulative flop count
of GB_AxB_flopcount (Source/GB_AxB_flopcount.c:32-37).  SURVEY.md 8(e) maps every shape of the path:

  matrix x matrix, saxpy   B's vectors are cut into N contiguous slices of equal flops, A (and M) are
                           replicated, every rank computes its slice of C; no data-path collective
  masked dot               M's ENTRIES are cut by owner vector (owner_aligned_mask): a rank takes the
                           pairs whose longer vector is B(:,j) with j in its range, or A(:,i) with i in
                           its range; A and B replicated; no data-path collective; the ranks' T are
                           disjoint parts of the whole T
  vector pull (A'*u)       A's vectors (= output entries) are cut by entry count; the slices of w are
                           all-gathered
  vector push (A*u)        u's entries are cut by flops; every rank gets a partial w over the whole
                           index range; the partial vectors are all-gathered and combined with the
                           monoid by one more multiply on the GPU: w = [w_0 .. w_{N-1}] * ones over the
                           semiring (add, FIRST) -- the same saxpy kernel, so the combination has the
                           reference's monoid semantics for every one of the built-in monoids
  scalars                  flop and entry counts are all-reduced (SUM)

Every rank passes the SAME host operands (replicated, as SURVEY.md 8e prescribes).  The local
multiply is injectable so that the partition / exchange logic is testable on a CPU-only box against
the oracle; the default is the CUDA library and there is no other production path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import (METHOD_DEFAULT, Matrix, Result, Semiring, TYPES, axb_host, partition_by_flops)

MASK_KEEP, MASK_DROP = 0x10000, 0x20000


@dataclass
class ShardedResult:
    local: Matrix               # this rank's slice of T (same dimensions as T; other vectors empty)
    lo: int                     # the slice: stored vectors / entries [lo, hi) of the sliced operand
    hi: int
    sliced: str                 # which operand was cut: "B", "M", "A" or "u"
    nnz: int                    # entries of the whole T (all ranks)
    flops: int                  # multiply-adds of the whole multiply (all ranks)
    mask_applied: bool
    method_used: int
    full: Optional[Matrix]      # the whole T, on every rank, if gather=True (always for vectors)
    local_info: dict


# ---------------------------------------------------------------------------------------------
# slicing and concatenation of the CSC-agnostic layout
# ---------------------------------------------------------------------------------------------
def slice_vectors(m: Matrix, lo: int, hi: int) -> Matrix:
    """Stored vectors [lo,hi) of m; dimensions unchanged.  Standard form keeps a full-length pointer
    array (the other vectors become empty); hypersparse form keeps only the slice's hyperlist."""
    s, e = int(m.p[lo]), int(m.p[hi])
    if m.h is not None:
        return Matrix(m.vlen, m.vdim, m.p[lo:hi + 1] - s, m.i[s:e], m.x[s:e], m.h[lo:hi], m.type)
    p = np.zeros(len(m.p), dtype=np.int64)
    p[lo:hi + 1] = m.p[lo:hi + 1] - s
    p[hi + 1:] = e - s
    return Matrix(m.vlen, m.vdim, p, m.i[s:e], m.x[s:e], None, m.type)


def slice_entries(u: Matrix, lo: int, hi: int) -> Matrix:
    """Entries [lo,hi) of an n-by-1 matrix (a GrB_Vector at the seam)."""
    return Matrix(u.vlen, 1, np.array([0, hi - lo], dtype=np.int64), u.i[lo:hi], u.x[lo:hi], None,
                  u.type)


def concat_slices(parts: list[Matrix]) -> Matrix:
    """Concatenate per-rank slices of T, given in rank order (vector ranges ascending)."""
    first = parts[0]
    i = np.concatenate([t.i for t in parts])
    x = np.concatenate([t.x for t in parts])
    if first.h is None:
        # every slice has a full-length cumulative pointer array that is flat outside its range
        p = np.sum([t.p for t in parts], axis=0).astype(np.int64)
        return Matrix(first.vlen, first.vdim, p, i, x, None, first.type)
    off, ps, hs = 0, [], []
    for t in parts:
        ps.append(t.p[:-1] + off)
        hs.append(t.h)
        off += int(t.p[-1])
    p = np.concatenate(ps + [np.array([off], dtype=np.int64)])
    return Matrix(first.vlen, first.vdim, p, i, x, np.concatenate(hs), first.type)


# ---------------------------------------------------------------------------------------------
# collectives (torch.distributed; tensors go through the GPU when the backend is NCCL)
# ---------------------------------------------------------------------------------------------
class _Comm:
    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.on else 0
        self.world = dist.get_world_size(group) if self.on else 1
        self.device = "cpu"
        if self.on and dist.get_backend(group) == "nccl":
            self.device = f"cuda:{torch.cuda.current_device()}"

    def sum_i64(self, values) -> list[int]:
        if not self.on:
            return [int(v) for v in values]
        t = self.torch.tensor([int(v) for v in values], dtype=self.torch.int64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return [int(v) for v in t.tolist()]

    def gather_i64(self, value: int) -> list[int]:
        if not self.on:
            return [int(value)]
        t = self.torch.zeros(self.world, dtype=self.torch.int64, device=self.device)
        t[self.rank] = int(value)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return [int(v) for v in t.tolist()]

    def allgather_bytes(self, a: np.ndarray, counts: list[int]) -> list[np.ndarray]:
        """all-gather of a variable-length 1-D array (counts[r] elements on rank r): padded to the
        longest, one all_gather_into_tensor, trimmed."""
        if not self.on:
            return [a]
        item = a.dtype.itemsize
        mx = max(max(counts), 1) * item
        src = self.torch.zeros(mx, dtype=self.torch.uint8, device=self.device)
        if a.size:
            raw = self.torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1))
            src[: raw.numel()] = raw.to(self.device)
        dst = self.torch.empty(mx * self.world, dtype=self.torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(dst, src, group=self.group)
        host = dst.cpu().numpy()
        return [host[r * mx: r * mx + counts[r] * item].copy().view(a.dtype) for r in range(self.world)]


# ---------------------------------------------------------------------------------------------
# balance weights
# ---------------------------------------------------------------------------------------------
def saxpy_flops_cum(M: Optional[Matrix], A: Matrix, B: Matrix) -> np.ndarray:
    """Cumulative Bflops of GB_AxB_flopcount (without the mask's range pruning: a balance weight,
    not a result) from host pointer arithmetic: Bflops[kk] = sum over B(k,j) of nnz(A(:,k))."""
    if A.h is None:
        lenA = np.diff(A.p)
        w = lenA[B.i]
    else:
        pos = np.searchsorted(A.h, B.i)
        pos[pos >= len(A.h)] = 0
        w = np.where(A.h[pos] == B.i, np.diff(A.p)[pos] if len(A.h) else 0, 0)
    cs = np.concatenate([[0], np.cumsum(w)]).astype(np.int64)
    return cs[B.p]


def _vec_lengths(m: Matrix, names: np.ndarray) -> np.ndarray:
    """length of the vectors of m called `names` (0 where absent)"""
    if m.h is None:
        return np.diff(m.p)[names]
    if len(m.h) == 0:
        return np.zeros(len(names), dtype=np.int64)
    pos = np.minimum(np.searchsorted(m.h, names), len(m.h) - 1)
    return np.where(m.h[pos] == names, np.diff(m.p)[pos], 0)


def masked_dot_walk_cum(M: Matrix, A: Matrix, B: Matrix) -> np.ndarray:
    """Per stored vector of M: sum over its entries (i,j) of min(len A(:,i), len B(:,j)) + 32 -- the
    length of the list the masked dot kernel walks for that entry -- cumulative."""
    names = M.h if M.h is not None else np.arange(M.vdim)
    lenB = np.repeat(_vec_lengths(B, names), np.diff(M.p))
    lenA = _vec_lengths(A, M.i)
    cs = np.concatenate([[0], np.cumsum(np.minimum(lenA, lenB) + 32)]).astype(np.int64)
    return cs[M.p]


class OwnerPartition:
    """The owner-aligned cut of a mask for C<M> = A'*B on N ranks.

    A rank computes the mask entries (i,j) whose LONGER vector ("owner", the one the masked dot kernel
    loads into shared memory) belongs to it: B(:,j) with j in the rank's range of M's vectors, or
    A(:,i) with i in the rank's range of row indices.  Cutting M by vectors alone would hand every
    rank 1/N of the pairs of each A-owned hub, so every rank would load every hub (measured on RMAT
    scale 22, N = 8: 4.0x instead of the 7x the kernels allow).  The parts are disjoint and cover M.
    Both ranges are balanced by a cost model of the kernels; rebalance() corrects the model with the
    ranks' measured times (set-up work: a calibration multiply per rank, outside any timed region)."""

    def __init__(self, M: Matrix, A: Matrix, B: Matrix, nparts: int, hub_len: int = 6144):
        self.M, self.nparts = M, nparts
        names = M.h if M.h is not None else np.arange(M.vdim, dtype=np.int64)
        cnt = np.diff(M.p)
        self.vpos = np.repeat(np.arange(M.nvec, dtype=np.int64), cnt)
        lenB = np.repeat(_vec_lengths(B, names), cnt)
        lenA = _vec_lengths(A, M.i)
        vlen = A.vlen
        self.b_owns = (lenB == vlen) | ((lenA != vlen) & (lenA <= lenB))   # the kernel's own rule
        # cost of a pair: the walked length + task set-up (worth ~32 probes); a probe of a hub owner
        # (loaded segment by segment, cursor per task) was measured at ~1.6x a probe of a regular
        # owner, and every segment of a hub is one more visit of the task (~24 probes' worth)
        olen = np.where(self.b_owns, lenB, lenA)
        is_hub = (olen > hub_len) & (olen != vlen)
        hub = np.where(is_hub, 8, 5)
        visits = np.where(is_hub, 24 * ((olen + hub_len - 1) // hub_len), 0)
        wB = np.where(self.b_owns, (lenA + visits + 32) * hub, 0).astype(np.float64)    # walk A(:,i)
        wA = np.where(~self.b_owns, (lenB + visits + 32) * hub, 0).astype(np.float64)
        # per vector of M (B-owned pairs) and per row index (A-owned pairs)
        self.wvec = np.add.reduceat(np.concatenate([wB, [0.0]]), np.minimum(M.p[:-1], len(wB))) * (cnt > 0) \
            if M.nvec else np.zeros(0)
        self.wrow = np.bincount(M.i, weights=wA, minlength=M.vlen)
        self.cut()

    def cut(self):
        """equal-cost contiguous ranges of vectors (jb) and of rows (ib)"""
        cumB = np.concatenate([[0], np.cumsum(np.rint(self.wvec))]).astype(np.int64)
        cumA = np.concatenate([[0], np.cumsum(np.rint(self.wrow))]).astype(np.int64)
        self.jb = partition_by_flops(cumB, self.nparts)
        self.ib = partition_by_flops(cumA, self.nparts)

    def ranges(self, part: int):
        return (int(self.jb[part]), int(self.jb[part + 1])), (int(self.ib[part]), int(self.ib[part + 1]))

    def mask(self, part: int) -> Matrix:
        """M_r: the entries of M that rank `part` computes (M's dimensions and vector list)"""
        M = self.M
        (jlo, jhi), (ilo, ihi) = self.ranges(part)
        sel = (self.b_owns & (self.vpos >= jlo) & (self.vpos < jhi)) \
            | (~self.b_owns & (M.i >= ilo) & (M.i < ihi))
        p = np.concatenate([[0], np.cumsum(np.bincount(self.vpos[sel], minlength=M.nvec))]).astype(np.int64)
        return Matrix(M.vlen, M.vdim, p, M.i[sel], M.x[sel], M.h, M.type)

    def rebalance(self, times) -> None:
        """times[r] = measured time of rank r's part under the current cut: the modelled cost of
        everything rank r owns is scaled by (its time / the mean), and the ranges are cut again."""
        t = np.asarray(times, dtype=np.float64)
        if len(t) != self.nparts or not np.all(t > 0):
            return
        f = t / t.mean()
        for r in range(self.nparts):
            (jlo, jhi), (ilo, ihi) = self.ranges(r)
            self.wvec[jlo:jhi] *= f[r]
            self.wrow[ilo:ihi] *= f[r]
        self.cut()


def owner_aligned_mask(M: Matrix, A: Matrix, B: Matrix, nparts: int, part: int, hub_len: int = 6144):
    """(M_r, (jlo, jhi), (ilo, ihi)) of OwnerPartition for one rank"""
    op = OwnerPartition(M, A, B, nparts, hub_len)
    rj, ri = op.ranges(part)
    return op.mask(part), rj, ri


def merge_disjoint(parts: list[Matrix]) -> Matrix:
    """Union of matrices with the same dimensions whose patterns are disjoint (the per-rank results of
    an owner-aligned masked dot): entries re-sorted inside every vector.  Hypersparse parts list only
    their own non-empty vectors, so vectors are matched by name."""
    first = parts[0]
    hyper = first.h is not None
    vname = [np.repeat(t.h if hyper else np.arange(t.nvec, dtype=np.int64), np.diff(t.p)) for t in parts]
    names = np.unique(np.concatenate([t.h for t in parts])) if hyper else None
    v = np.concatenate(vname)
    i = np.concatenate([t.i for t in parts])
    x = np.concatenate([t.x for t in parts])
    vpos = np.searchsorted(names, v) if hyper else v
    nvec = len(names) if hyper else first.nvec
    order = np.lexsort((i, vpos))
    p = np.concatenate([[0], np.cumsum(np.bincount(vpos, minlength=nvec))]).astype(np.int64)
    return Matrix(first.vlen, first.vdim, p, i[order], x[order], names, first.type)


# ---------------------------------------------------------------------------------------------
# the multiply
# ---------------------------------------------------------------------------------------------
def _is_vector(m: Matrix) -> bool:
    return m.vdim == 1 and m.h is None


def mxm(M: Optional[Matrix], mask_comp: bool, A: Matrix, B: Matrix, semiring: Semiring,
        do_adotb: bool = False, method: int = METHOD_DEFAULT, *, gather: bool = False, group=None,
        multiply: Optional[Callable[..., Result]] = None) -> ShardedResult:
    """C<M> = A*B (or A'*B) at the GB_AxB_parallel seam on all ranks of `group`.  Every rank passes
    the same host operands and gets its slice of T (plus the whole T if gather=True; vector results
    are always whole)."""
    comm = _Comm(group)
    mul = multiply or axb_host
    if _is_vector(B) and B.nnz > 0 and A.nnz > 0 and (M is None or _is_vector(M)):
        if do_adotb:
            return _vector_pull(comm, mul, M, mask_comp, A, B, semiring, method)
        return _vector_push(comm, mul, M, mask_comp, A, B, semiring, method)

    W, r = comm.world, comm.rank
    if do_adotb and M is not None and not mask_comp:
        # masked dot: every mask entry is an independent dot product; a rank takes the entries whose
        # owner vector is its own (owner_aligned_mask).  Its T is a disjoint part of the whole T
        # (not a contiguous run of vectors): gather=True merges the parts.
        Mr, (lo, hi), _ = owner_aligned_mask(M, A, B, W, r)
        res = mul(Mr, False, A, B, semiring, True, method)
        nnz_all, flops_all = comm.sum_i64([res.info["nnz"], res.info["flops"]])
        T = res.matrix
        full = None
        if gather:
            full = T if W == 1 else merge_disjoint(_gather_parts(comm, T))
        return ShardedResult(T, lo, hi, "M(owner)", nnz_all, flops_all, True,
                             int(res.info["method_used"]), full, res.info)
    cum, sliced, what = saxpy_flops_cum(M, A, B), B, "B"
    if do_adotb:
        # unmasked / complemented dot: every (i,j) pair is computed; cost of B(:,j) ~ its length
        cum = (B.p + np.arange(len(B.p))).astype(np.int64)
    bounds = partition_by_flops(cum, W)
    lo, hi = int(bounds[r]), int(bounds[r + 1])
    mine = slice_vectors(sliced, lo, hi)
    Mr, Br = M, mine
    # the saxpy mask rule (GB_AxB_sequential.c:88-95) compares GLOBAL counts: decide once.  The flop
    # count that the rule uses honours the mask's range pruning, so it comes from the multiplies
    # themselves: first pass with KEEP, and if the global rule says "drop", redo without the mask.
    flags = 0
    if M is not None and not mask_comp and not do_adotb:
        flags = MASK_KEEP
    res = mul(Mr, mask_comp, A, Br, semiring, do_adotb, method | flags)
    nnz_all, flops_all = comm.sum_i64([res.info["nnz"], res.info["flops"]])
    if flags == MASK_KEEP and flops_all <= M.nnz:
        res = mul(None, False, A, Br, semiring, do_adotb, method | MASK_DROP)
        nnz_all, flops_all = comm.sum_i64([res.info["nnz"], res.info["flops"]])
    T = res.matrix
    full = None
    if gather:
        full = _gather_matrix(comm, T)
    return ShardedResult(T, lo, hi, what, nnz_all, flops_all, bool(res.info["mask_applied"]),
                         int(res.info["method_used"]), full, res.info)


def _gather_parts(comm: _Comm, T: Matrix) -> list:
    """every rank's T (same dimensions and vector list), on every rank"""
    nz = comm.gather_i64(T.nnz)
    Is = comm.allgather_bytes(T.i, nz)
    Xs = comm.allgather_bytes(T.x, nz)
    nv = comm.gather_i64(len(T.p))
    Ps = comm.allgather_bytes(T.p, nv)
    Hs = [None] * comm.world
    if T.h is not None:
        Hs = comm.allgather_bytes(T.h, [n - 1 for n in nv])
    return [Matrix(T.vlen, T.vdim, Ps[q], Is[q], Xs[q], Hs[q], T.type) for q in range(comm.world)]


def _gather_matrix(comm: _Comm, T: Matrix) -> Matrix:
    if comm.world == 1:
        return T
    nz = comm.gather_i64(T.nnz)
    Is = comm.allgather_bytes(T.i, nz)
    Xs = comm.allgather_bytes(T.x, nz)
    nv = comm.gather_i64(len(T.p))
    Ps = comm.allgather_bytes(T.p, nv)
    Hs = [None] * comm.world
    if T.h is not None:
        Hs = comm.allgather_bytes(T.h, [n - 1 for n in nv])
    parts = [Matrix(T.vlen, T.vdim, Ps[q], Is[q], Xs[q], Hs[q], T.type) for q in range(comm.world)]
    return concat_slices(parts)


def _vector_pull(comm, mul, M, mask_comp, A, u, semiring, method) -> ShardedResult:
    """w<M> = A'*u: rank r owns a contiguous block of A's vectors, i.e. of w's entries."""
    W, r = comm.world, comm.rank
    cum = (A.p + np.arange(len(A.p))).astype(np.int64)      # entries + one unit per vector
    bounds = partition_by_flops(cum, W)
    lo, hi = int(bounds[r]), int(bounds[r + 1])
    res = mul(M, mask_comp, slice_vectors(A, lo, hi), u, semiring, True, method)
    T = res.matrix
    nnz_all, flops_all = comm.sum_i64([T.nnz, res.info["flops"]])
    nz = comm.gather_i64(T.nnz)
    Is = comm.allgather_bytes(T.i, nz)
    Xs = comm.allgather_bytes(T.x, nz)
    i, x = np.concatenate(Is), np.concatenate(Xs)
    full = Matrix(T.vlen, 1, np.array([0, len(i)], dtype=np.int64), i, x, None, T.type)
    return ShardedResult(T, lo, hi, "A", nnz_all, flops_all, bool(res.info["mask_applied"]),
                         int(res.info["method_used"]), full, res.info)


def _vector_push(comm, mul, M, mask_comp, A, u, semiring, method) -> ShardedResult:
    """w<M> = A*u: rank r expands a slice of u's entries into a partial w; the partial vectors are
    all-gathered and combined by w = [w_0 .. w_{N-1}] * ones over (add, FIRST) on the GPU."""
    W, r = comm.world, comm.rank
    lens = _vec_lengths(A, u.i)
    cum = np.concatenate([[0], np.cumsum(lens + 1)]).astype(np.int64)
    bounds = partition_by_flops(cum, W)
    lo, hi = int(bounds[r]), int(bounds[r + 1])
    flags = MASK_KEEP if (M is not None and not mask_comp) else 0
    ur = slice_entries(u, lo, hi)
    if hi > lo:
        res = mul(M, mask_comp, A, ur, semiring, False, method | flags)
        T, info = res.matrix, res.info
    else:
        # an empty slice of u: nothing to expand (the seam is never entered with an empty vector)
        ztype = semiring.ztype
        T = Matrix(A.vlen, 1, np.zeros(2, np.int64), np.zeros(0, np.int64),
                   np.zeros(0, TYPES[ztype][1]), None, ztype)
        info = {"nnz": 0, "flops": 0, "mask_applied": 1 if M is not None else 0,
                "method_used": 1001, "nvec": 1}
    _, flops_all, applied_all = comm.sum_i64([0, info["flops"], 1 if info["mask_applied"] else 0])
    if flags == MASK_KEEP and flops_all <= M.nnz:
        # the global rule drops the mask: redo the expansion without it
        if hi > lo:
            res = mul(None, False, A, ur, semiring, False, method | MASK_DROP)
            T, info = res.matrix, res.info
        _, flops_all, applied_all = comm.sum_i64([0, info["flops"], 0])
    mask_applied = (M is not None) and applied_all > 0
    if W == 1:
        full = T
    else:
        nz = comm.gather_i64(T.nnz)
        Is = comm.allgather_bytes(T.i, nz)
        Xs = comm.allgather_bytes(T.x, nz)
        P = Matrix(T.vlen, W, np.concatenate([[0], np.cumsum(nz)]).astype(np.int64),
                   np.concatenate(Is), np.concatenate(Xs), None, T.type)
        if P.nnz == 0:
            full = T
        else:
            ones = Matrix(W, 1, np.array([0, W], dtype=np.int64), np.arange(W, dtype=np.int64),
                          np.ones(W, dtype=TYPES[T.type][1]), None, T.type)
            combine = Semiring(semiring.add, "FIRST", T.type)
            full = mul(None, False, P, ones, combine, False, METHOD_DEFAULT).matrix
    return ShardedResult(T, lo, hi, "u", full.nnz, flops_all, mask_applied,
                         int(info["method_used"]), full, info)
