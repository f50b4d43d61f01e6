"""cache_check.py -- run in its own process by tests/test_gpu_cache.py (GraphBLAS is started once per
process, here with GxB_init (gb200_host_*) so that every free / realloc of a GraphBLAS array reaches the
library's allocator): the operand residency cache behind the shim must serve repeated multiplies from
HBM and must never serve a stale copy -- after GrB_Matrix_setElement on an existing entry (in place),
after a new entry (pending tuple -> GB_wait rebuilds the arrays), and after an operand is freed and a
new one is imported (the allocator recycles the same page-locked blocks)."""
import os
import sys

import numpy as np

import gen
import grbref
from parity import compare, export_csr, import_sp


def mxm(G, a, b, gpu, n, desc=None):
    c = G.matrix_new("FP64", n, n)
    G.use_gpu(gpu)
    try:
        G.mxm(c, None, None, "GxB_PLUS_TIMES_FP64", a, b, desc)
        G.matrix_nvals(c)
    finally:
        G.use_gpu(False)
    return export_csr(G, c)


def check(G, a, b, n, what, desc=None):
    ref = mxm(G, a, b, False, n, desc)
    got = mxm(G, a, b, True, n, desc)
    ok, why = compare(ref, got, "PLUS")
    assert ok, f"{what}: {why}"


def main():
    G = grbref.GraphBLAS.get(with_shim=True, pinned=True)
    n = int(os.environ.get("GB200_CACHE_CHECK_N", "3000"))
    A = gen.er(n, n, 50 * n, 11)
    B = gen.er(n, n, 50 * n, 12)
    a, b = import_sp(G, A, "FP64", "CSR"), import_sp(G, B, "FP64", "CSR")
    s0 = G.shim_cache(True)
    check(G, a, b, n, "first multiply")
    s1 = G.shim_cache()
    assert s1["misses"] - s0["misses"] == 2 and s1["hits"] == s0["hits"], (s0, s1)
    check(G, a, b, n, "second multiply")
    s2 = G.shim_cache()
    assert s2["hits"] - s1["hits"] == 2 and s2["misses"] == s1["misses"], (s1, s2)
    # an existing entry is overwritten in place: no free, no realloc -- only the interposed writer sees it
    i, j = int(A.nonzero()[0][7]), int(A.nonzero()[1][7])
    G.matrix_set_element(a, "FP64", i, j, 1234.5)
    check(G, a, b, n, "after setElement on an existing entry")
    s3 = G.shim_cache()
    assert s3["invalidations"] > s2["invalidations"], (s2, s3)
    # a new entry: a pending tuple, assembled by GB_wait inside the next multiply
    free = np.setdiff1d(np.arange(n), A[5].indices)
    G.matrix_set_element(a, "FP64", 5, int(free[0]), -7.25)
    check(G, a, b, n, "after setElement of a new entry")
    check(G, a, b, n, "and again, from the cache")
    s4 = G.shim_cache()
    assert s4["hits"] > s3["hits"], (s3, s4)
    # free an operand and import another one of the same shape: the pinned allocator hands out the same blocks
    G.matrix_free(b)
    B2 = gen.er(n, n, 50 * n, 13)
    B2 = B2.tocsr()[:, :]
    b = import_sp(G, B2, "FP64", "CSR")
    check(G, a, b, n, "after an operand was freed and replaced")
    # T stays resident: GrB_reduce right after the multiply starts from HBM (no upload of T)
    c = G.matrix_new("FP64", n, n)
    G.use_gpu(True)
    try:
        G.mxm(c, None, None, "GxB_PLUS_TIMES_FP64", a, b, None)
        s5 = G.shim_cache()
        got = G.matrix_reduce(c, "FP64", "GxB_MAX_FP64_MONOID")
        s6 = G.shim_cache()
    finally:
        G.use_gpu(False)
    want = G.matrix_reduce(c, "FP64", "GxB_MAX_FP64_MONOID")
    assert got == want, (got, want)
    assert s6["hits"] - s5["hits"] == 1 and s6["misses"] == s5["misses"], (s5, s6)
    # transposed operands (row f2): GB_AxB_meta's transposes run on the device (interposed GB_transpose),
    # from the resident copy of A; A' is adopted, so the multiply finds it in HBM; the reference frees A'
    # after the multiply (the entry goes with it), and a change of A must show in the next A'
    G.shim_transpose_min(0)
    t0, s7 = G.shim_transpose_calls(), G.shim_cache()
    descs = [G.descriptor(inp0=grbref.GrB_TRAN, method=grbref.GxB_AxB_GUSTAVSON),
             G.descriptor(inp1=grbref.GrB_TRAN, method=grbref.GxB_AxB_GUSTAVSON),
             G.descriptor(inp0=grbref.GrB_TRAN, inp1=grbref.GrB_TRAN, method=grbref.GxB_AxB_GUSTAVSON)]
    for k, d in enumerate(descs):
        check(G, a, b, n, f"transposed operands, descriptor {k}", d)
    t1, s8 = G.shim_transpose_calls(), G.shim_cache()
    assert t1 - t0 >= 3, (t0, t1)
    assert s8["hits"] - s7["hits"] >= 2 * (t1 - t0), (s7, s8, t1 - t0)     # A for the transpose, A' for the multiply
    G.matrix_set_element(a, "FP64", i, j, -99.5)
    G.matrix_set_element(b, "FP64", int(B2.nonzero()[0][3]), int(B2.nonzero()[1][3]), 17.25)
    for k, d in enumerate(descs):
        check(G, a, b, n, f"transposed operands after setElement, descriptor {k}", d)
    G.shim_transpose_min(65536)
    G.shim_cache(False)
    print("cache_check: ok", G.shim_cache(), "device transposes", G.shim_transpose_calls() - t0)


if __name__ == "__main__":
    sys.exit(main())
