"""The N>1 path on CPU: graphblas_b200.sharded over torch.distributed (gloo, world_size 2 and 3) with
the oracle injected as the local multiply, so that the partitioning / exchange / concatenation logic
is checked without a GPU; one world_size-2 run uses the EMULATED library (tests/emulated.py: the product's
own gb200_AxB_host compiled for the host) as the local multiply instead.  The GPU legs of the same cases are
in test_gpu_seam.py (N=1, the CUDA library as the local multiply) and bench.py --gpus N (NCCL)."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import gen
import graphblas_b200 as gb
import oracle_c
from graphblas_b200 import sharded

INT = np.int64


def oracle_multiply(M, mask_comp, A, B, semiring, do_adotb=False, method=0, **_):
    """Stand-in for gb.axb_host with the reference's semantics, including the mask-policy flags of
    include/gb_b200.h (KEEP / DROP decide the saxpy mask rule for all slices at once)."""
    keep, drop = bool(method & sharded.MASK_KEEP), bool(method & sharded.MASK_DROP)
    if drop:
        M = None
    info = {}
    T = oracle_c.axb(M, mask_comp, A, B, semiring, do_adotb, info=info)
    flops = 0
    if not do_adotb:
        use_m = M if (M is not None and not mask_comp) else None
        flops = int(oracle_c.flopcount(use_m, A, B)[1])
        if use_m is not None and not info["mask_applied"]:
            flops = int(oracle_c.flopcount(None, A, B)[1])
            if keep:
                # the slice's own rule dropped the mask but the caller wants it kept: filter T by the
                # (valued) mask -- identical to the masked kernel for the exact semirings used here
                T = _apply_mask(T, M)
                info["mask_applied"] = 1
                flops = int(oracle_c.flopcount(M, A, B)[1])
    info.update(nnz=T.nnz, flops=flops, nvec=T.nvec)
    return gb.Result(T, info)


def _apply_mask(T: gb.Matrix, M: gb.Matrix) -> gb.Matrix:
    assert T.h is None and M.h is None
    keep = np.zeros(T.nnz, dtype=bool)
    for j in range(T.vdim):
        t0, t1, m0, m1 = T.p[j], T.p[j + 1], M.p[j], M.p[j + 1]
        mi = M.i[m0:m1][M.x[m0:m1] != 0]
        keep[t0:t1] = np.isin(T.i[t0:t1], mi)
    cnt = np.add.reduceat(np.concatenate([keep, [False]]).astype(np.int64),
                          np.minimum(T.p[:-1], T.nnz)) if T.nnz else np.zeros(T.vdim, np.int64)
    cnt[np.diff(T.p) == 0] = 0
    p = np.concatenate([[0], np.cumsum(cnt)])
    return gb.Matrix(T.vlen, T.vdim, p, T.i[keep], T.x[keep], None, T.type)


def same(ref: gb.Matrix, got: gb.Matrix, what: str):
    assert (ref.vlen, ref.vdim) == (got.vlen, got.vdim), what
    assert (ref.h is None) == (got.h is None), what + ": hypersparsity"
    assert np.array_equal(ref.p, got.p), what + ": pointers"
    if ref.h is not None:
        assert np.array_equal(ref.h, got.h), what + ": hyperlist"
    assert np.array_equal(ref.i, got.i), what + ": pattern"
    assert ref.type == got.type and np.array_equal(ref.x, got.x), what + ": values"


def cases():
    """(name, M, mask_comp, A, B, semiring, do_adotb) -- integer / bool / MIN semirings: bit-exact"""
    n = 90
    A = gb.Matrix.from_scipy(gen.er(n, n, 9 * n, 11, INT, lo=1, hi=6).tocsc())
    B = gb.Matrix.from_scipy(gen.er(n, 70, 8 * n, 12, INT, lo=1, hi=6).tocsc())
    Bsq = gb.Matrix.from_scipy(gen.er(n, n, 8 * n, 16, INT, lo=1, hi=6).tocsc())
    M = gb.Matrix.from_scipy(gen.er(n, 70, 6 * n, 13, np.bool_).tocsc())
    Mdense = gb.Matrix.from_scipy(gen.er(n, 70, 60 * n, 14, np.bool_).tocsc())
    Msq = gb.Matrix.from_scipy(gen.er(n, n, 7 * n, 15, np.bool_).tocsc())
    pt = gb.Semiring("PLUS", "TIMES", "INT64")
    out = [
        ("saxpy", None, False, A, B, pt, False),
        ("saxpy masked", M, False, A, B, pt, False),
        ("saxpy mask dropped by the global rule", Mdense, False, A, B, pt, False),
        ("saxpy complemented mask", M, True, A, B, pt, False),
        ("saxpy hyper B", None, False, A, B.to_hyper(), pt, False),
        ("saxpy hyper A and M", M.to_hyper(), False, A.to_hyper(), B, pt, False),
        ("dot masked", Msq, False, A, Bsq, pt, True),
        ("dot masked hyper M", Msq.to_hyper(), False, A, Bsq, pt, True),
        ("dot complemented mask", Msq, True, A, Bsq, pt, True),
        ("dot", None, False, A, Bsq, pt, True),
    ]
    # vectors: n-by-1 operands (GrB_mxv / GrB_vxm at the seam)
    Ab = gb.Matrix.from_scipy(gen.er(n, n, 6 * n, 21, np.bool_).tocsc())
    q = gb.Matrix.from_scipy(gen.er(n, 1, 12, 22, np.bool_).tocsc())
    v = gb.Matrix.from_scipy(gen.er(n, 1, 30, 23, np.bool_).tocsc())
    ll = gb.Semiring("LOR", "LAND", "BOOL")
    Af = gb.Matrix.from_scipy(gen.er(n, n, 6 * n, 24, np.float64, lo=1, hi=9).tocsc())
    d = gb.Matrix(n, 1, np.array([0, n]), np.arange(n), np.arange(n, dtype=np.float64) + 1.0, None, "FP64")
    ds = gb.Matrix.from_scipy(gen.er(n, 1, 25, 25, np.float64, lo=1, hi=9).tocsc())
    mp = gb.Semiring("MIN", "PLUS", "FP64")
    out += [
        ("vector push, complemented mask (BFS)", v, True, Ab, q, ll, False),
        ("vector push, mask", v, False, Ab, q, ll, False),
        ("vector push, no mask, MIN_PLUS", None, False, Af, ds, mp, False),
        ("vector pull, dense vector (SSSP)", None, False, Af, d, mp, True),
        ("vector pull, complemented mask", v, True, Ab, q, ll, True),
        ("vector push hyper A", None, False, Af.to_hyper(), ds, mp, False),
    ]
    return out


# the cases run with the EMULATED library as the local multiply (tests/emulated.py): the product's own
# gb200_AxB_host, mask-policy flags included, instead of the oracle stand-in above
# (not the push with a complemented mask: the library applies <!M> inside the kernel and reports
# mask_applied, where the reference's saxpy ignores it -- same C, different T; test_gpu_seam.py covers it)
EMULATED_CASES = ("saxpy masked", "saxpy mask dropped by the global rule", "saxpy hyper A and M", "dot masked",
                  "vector push, no mask, MIN_PLUS", "vector pull, dense vector (SSSP)",
                  "vector pull, complemented mask")


def _worker(rank: int, world: int, port: int, errq, emulated_lib: bool = False):
    try:
        import contextlib
        import torch.distributed as dist
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                                world_size=world)
        if emulated_lib:
            import emulated
        for name, M, comp, A, B, sr, dot in cases():
            if emulated_lib and name not in EMULATED_CASES:
                continue
            ref_info = {}
            ref = oracle_c.axb(M, comp, A, B, sr, dot, info=ref_info)
            with (emulated.swapped() if emulated_lib else contextlib.nullcontext()):
                r = sharded.mxm(M, comp, A, B, sr, dot, gather=True,
                                multiply=None if emulated_lib else oracle_multiply)
            what = f"[{world} ranks, rank {rank}] {name}"
            assert r.full is not None, what
            got = r.full
            if got.h is not None and ref.h is None:
                got = got.to_standard()
            same(ref, got, what)
            assert r.nnz == ref.nnz, what + ": global nnz"
            assert r.mask_applied == bool(ref_info["mask_applied"]), what + ": mask_applied"
            assert 0 <= r.lo <= r.hi, what
            if r.sliced == "B" and world > 1:
                # slices are disjoint and cover: the local nnz add up (checked through r.nnz) and the
                # local T has entries only inside its own range of vectors
                cnt = np.diff(r.local.p)
                if r.local.h is None:
                    assert cnt[:r.lo].sum() == 0 and cnt[r.hi:].sum() == 0, what + ": slice range"
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put(f"rank {rank}:\n{traceback.format_exc()}")
        raise


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,emulated_lib", [(2, False), (3, False), (2, True)])
def test_sharded_multiply_gloo(world, emulated_lib):
    import torch.multiprocessing as mp
    if emulated_lib:
        import emulated
        emulated.library()              # built here once; the ranks load it from the cache
    ctx = mp.get_context("spawn")
    errq = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, errq, emulated_lib)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    msgs = []
    while not errq.empty():
        msgs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.kill()
            msgs.append("a rank hung")
    assert not msgs and all(p.exitcode == 0 for p in procs), "\n".join(msgs)


def test_sharded_single_process_matches_oracle():
    """world_size 1 (no process group): the sharded entry point is the plain multiply"""
    for name, M, comp, A, B, sr, dot in cases():
        ref = oracle_c.axb(M, comp, A, B, sr, dot)
        r = sharded.mxm(M, comp, A, B, sr, dot, gather=True, multiply=oracle_multiply)
        got = r.full
        if got.h is not None and ref.h is None:
            got = got.to_standard()
        same(ref, got, name)


def test_partition_is_flop_balanced():
    """equal-flop contiguous slices (GB_AxB_flopcount.c:32-37): no slice exceeds the ideal share by
    more than one vector's flops"""
    A = gb.Matrix.from_scipy(gen.rmat_scipy(10, 8).tocsc().astype(np.int64))
    cum = sharded.saxpy_flops_cum(None, A, A)
    ref_cum, total = oracle_c.flopcount(None, A, A)
    assert np.array_equal(cum, ref_cum) and cum[-1] == total
    for parts in (2, 4, 8):
        b = gb.partition_by_flops(cum, parts)
        assert b[0] == 0 and b[-1] == A.nvec and np.all(np.diff(b) >= 0)
        per = np.diff(cum[b])
        assert per.sum() == total
        assert per.max() <= total / parts + np.diff(cum).max()


def test_owner_aligned_mask_parts_are_disjoint_and_cover():
    """every mask entry goes to exactly one rank, and an owner vector never serves two ranks"""
    S = gen.rmat_scipy(11, 8).tocsc().astype(np.int64)
    L = gb.Matrix.from_scipy(__import__("scipy.sparse", fromlist=["tril"]).tril(S, -1).tocsc())
    U = gb.Matrix.from_scipy(__import__("scipy.sparse", fromlist=["triu"]).triu(S, 1).tocsc())
    for M in (L, L.to_hyper()):
        for W in (2, 5):
            parts = [sharded.owner_aligned_mask(M, U, L, W, r) for r in range(W)]
            assert sum(p[0].nnz for p in parts) == M.nnz
            merged = sharded.merge_disjoint([p[0] for p in parts])
            assert np.array_equal(merged.p, M.p) and np.array_equal(merged.i, M.i)
            lenA, lenB = np.diff(U.p), np.diff(L.p)
            names = M.h if M.h is not None else np.arange(M.vdim)
            for Mr, (jlo, jhi), (ilo, ihi) in parts:
                j = names[np.repeat(np.arange(Mr.nvec), np.diff(Mr.p))]
                b_owns = lenA[Mr.i] <= lenB[j]
                pos = np.searchsorted(names, j)
                assert np.all((pos[b_owns] >= jlo) & (pos[b_owns] < jhi))
                assert np.all((Mr.i[~b_owns] >= ilo) & (Mr.i[~b_owns] < ihi))


def test_owner_partition_rebalance_converges():
    """rebalance() with measured times: a cost the model does not know (one half of the index range
    50 % dearer) is learnt from the ranks' times and the parts even out"""
    S = gen.rmat_scipy(12, 8).tocsc().astype(np.int64)
    import scipy.sparse as sps
    L = gb.Matrix.from_scipy(sps.tril(S, -1).tocsc())
    U = gb.Matrix.from_scipy(sps.triu(S, 1).tocsc())
    W = 4
    op = sharded.OwnerPartition(L, U, L, W)
    true_vec = op.wvec * np.where(np.arange(L.nvec) < L.nvec // 2, 1.0, 1.5)
    true_row = op.wrow * np.where(np.arange(L.vlen) < L.vlen // 2, 1.0, 1.5)

    def times():
        out = []
        for r in range(W):
            (jlo, jhi), (ilo, ihi) = op.ranges(r)
            out.append(true_vec[jlo:jhi].sum() + true_row[ilo:ihi].sum())
        return np.array(out)
    t0 = times()
    for _ in range(3):
        op.rebalance(times())
    t1 = times()
    assert t1.max() / t1.mean() < t0.max() / t0.mean()
    assert t1.max() / t1.mean() < 1.05
    # the parts still cover the mask exactly once
    assert sum(op.mask(r).nnz for r in range(W)) == L.nnz
