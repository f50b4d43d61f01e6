"""GPU parity tests: the B200 path (through the interposed GB_AxB_parallel, i.e. through the C ABI of
libgb_b200.so) against the compiled reference on the same inputs.  The cases follow the reference's
own mxm tests (Test/test06.m: semirings x methods x transposes x mask; test74/75: dot; test88:
hypersparse + heap; test20: mxv/vxm + accum) at sizes the CPU reference finishes in milliseconds."""
import numpy as np
import pytest
import scipy.sparse as sp

import gen
import grbref
from grbref import (GxB_DEFAULT, GrB_REPLACE, GrB_SCMP, GrB_TRAN, GxB_AxB_GUSTAVSON, GxB_AxB_HEAP,
                    GxB_AxB_DOT)
from parity import check_mxm, check_mv, REF_ONLY

pytestmark = pytest.mark.gpu

NP = grbref.NP_OF


# ---------------------------------------------------------------------------------------------
# config 1 shape: C = A*B, PLUS_TIMES_FP64, ER, no mask (reference picks Gustavson)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt", ["CSR", "CSC"])
def test_cfg1_er_plus_times_fp64(G, fmt):
    n = 4096
    A = gen.er(n, n, 8 * n, 1)
    B = gen.er(n, n, 8 * n, 2)
    ref, got = check_mxm(G, A=A, B=B, type_="FP64", semiring="GxB_PLUS_TIMES_FP64", fmt=fmt)
    assert ref["nvals"] > 0


@pytest.mark.parametrize("semiring,type_", [
    ("GxB_PLUS_TIMES_INT64", "INT64"), ("GxB_PLUS_TIMES_UINT32", "UINT32"),
    ("GxB_MIN_PLUS_FP64", "FP64"), ("GxB_MAX_MIN_INT32", "INT32"),
    ("GxB_LOR_LAND_BOOL", "BOOL"), ("GxB_PLUS_TIMES_FP32", "FP32"),
    ("GxB_TIMES_PLUS_INT8", "INT8"), ("GxB_MIN_MAX_UINT16", "UINT16"),
    ("GxB_LXOR_LOR_BOOL", "BOOL"), ("GxB_EQ_LT_FP32", "FP32"), ("GxB_PLUS_DIV_INT16", "INT16"),
    ("GxB_PLUS_MINUS_FP64", "FP64"), ("GxB_MAX_FIRST_UINT8", "UINT8"),
    ("GxB_LAND_GE_INT64", "INT64"), ("GxB_TIMES_SECOND_FP64", "FP64"),
    ("GxB_PLUS_ISGT_UINT64", "UINT64"), ("GxB_MIN_DIV_INT32", "INT32"),
])
@pytest.mark.parametrize("fmt", ["CSR", "CSC"])
def test_unmasked_semirings(G, semiring, type_, fmt):
    A = gen.er(300, 200, 2400, 3, NP[type_])
    B = gen.er(200, 250, 2000, 4, NP[type_])
    check_mxm(G, A=A, B=B, type_=type_, semiring=semiring, fmt=fmt)


@pytest.mark.parametrize("inp0,inp1", [(GxB_DEFAULT, GxB_DEFAULT), (GrB_TRAN, GxB_DEFAULT),
                                       (GxB_DEFAULT, GrB_TRAN), (GrB_TRAN, GrB_TRAN)])
@pytest.mark.parametrize("method", [GxB_DEFAULT, GxB_AxB_GUSTAVSON, GxB_AxB_HEAP, GxB_AxB_DOT])
@pytest.mark.parametrize("fmt", ["CSR", "CSC"])
def test_transposes_and_methods(G, inp0, inp1, method, fmt):
    """the 16 CSR/CSC x transpose cases folded by GB_AxB_meta.c:97-180, under every method request"""
    A = gen.er(120, 120, 900, 5, np.int32)
    B = gen.er(120, 120, 800, 6, np.int32)
    check_mxm(G, A=A, B=B, type_="INT32", semiring="GxB_PLUS_MINUS_INT32", fmt=fmt, inp0=inp0,
              inp1=inp1, method=method)


@pytest.mark.parametrize("maskd", [GxB_DEFAULT, GrB_SCMP])
@pytest.mark.parametrize("outp", [GxB_DEFAULT, GrB_REPLACE])
@pytest.mark.parametrize("method", [GxB_DEFAULT, GxB_AxB_GUSTAVSON, GxB_AxB_DOT])
@pytest.mark.parametrize("inp0", [GxB_DEFAULT, GrB_TRAN])
def test_masked(G, maskd, outp, method, inp0):
    """valued masks (entries that are present but false do not admit), complement, replace, accum"""
    n = 150
    A = gen.er(n, n, 1500, 7, np.float64)
    B = gen.er(n, n, 1400, 8, np.float64)
    M = gen.er(n, n, 3000, 9, np.int8, lo=0, hi=2)          # about half the entries are zero
    Cinit = gen.er(n, n, 700, 10, np.float64)
    check_mxm(G, A=A, B=B, type_="FP64", semiring="GxB_PLUS_TIMES_FP64", M=M, mtype="INT8",
              Cinit=Cinit, accum="GrB_PLUS_FP64", mask=maskd, outp=outp, method=method, inp0=inp0)


def test_mask_dense_is_dropped(G):
    """flops <= nnz(M): the saxpy path discards the mask (GB_AxB_sequential.c:88-95); C is the same"""
    n = 60
    A = gen.er(n, n, 70, 11, np.int64)
    B = gen.er(n, n, 70, 12, np.int64)
    M = sp.csr_matrix(np.ones((n, n), dtype=np.bool_))
    check_mxm(G, A=A, B=B, type_="INT64", semiring="GxB_PLUS_TIMES_INT64", M=M, mtype="BOOL")


@pytest.mark.parametrize("fmt", ["HyperCSR", "HyperCSC"])
@pytest.mark.parametrize("method", [GxB_DEFAULT, GxB_AxB_DOT])
@pytest.mark.parametrize("masked", [False, True])
def test_hypersparse(G, fmt, method, masked):
    """hypersparse operands (Test/test88.m): vectors found through the hyperlist, hypersparse T"""
    n = 5000
    A = gen.er(n, n, 600, 13, np.int32)
    B = gen.er(n, n, 500, 14, np.int32)
    # give them some common structure so that products exist
    A = (A + sp.csr_matrix((np.ones(40, np.int32), (np.arange(40) * 7, np.arange(40) * 11)), shape=(n, n))).tocsr()
    B = (B + sp.csr_matrix((np.ones(40, np.int32), (np.arange(40) * 11, np.arange(40) * 3)), shape=(n, n))).tocsr()
    M = gen.er(n, n, 4000, 15, np.bool_) + sp.csr_matrix(
        (np.ones(40, np.bool_), (np.arange(40) * 7, np.arange(40) * 3)), shape=(n, n)) if masked else None
    check_mxm(G, A=A, B=B, type_="INT32", semiring="GxB_PLUS_TIMES_INT32", fmt=fmt, method=method,
              M=M.tocsr() if masked else None, mtype="BOOL")


def test_heavy_vectors(G):
    """a few vectors whose flop count exceeds every shared-memory bin (bitmap path), plus every
    lighter bin in the same multiply"""
    n = 3000
    rng = np.random.default_rng(16)
    A = gen.er(n, n, 40 * n, 17, np.int64, lo=1, hi=4).tolil()
    B = gen.er(n, n, 4 * n, 18, np.int64, lo=1, hi=4).tolil()
    for r in (5, 77, 1500):                                  # dense rows of B -> heavy (CSR: row-wise)
        cols = rng.choice(n, 1200, replace=False)
        B[r, cols] = 1
    for r in (9, 800):                                       # medium rows
        cols = rng.choice(n, 150, replace=False)
        B[r, cols] = 2
    A, B = A.tocsr(), B.tocsr()
    check_mxm(G, A=B, B=A, type_="INT64", semiring="GxB_PLUS_TIMES_INT64")
    M = gen.er(n, n, 30 * n, 19, np.bool_)
    check_mxm(G, A=B, B=A, type_="INT64", semiring="GxB_PLUS_TIMES_INT64", M=M, mtype="BOOL")
    check_mxm(G, A=B, B=A, type_="INT64", semiring="GxB_MIN_PLUS_INT64", fmt="CSC")


def test_empty_and_ragged(G):
    """empty operands, empty result, 1 x n and n x 1 shapes"""
    Z = sp.csr_matrix((50, 40), dtype=np.float64)
    B = gen.er(40, 30, 100, 20, np.float64)
    check_mxm(G, A=Z, B=B, type_="FP64", semiring="GxB_PLUS_TIMES_FP64")
    check_mxm(G, A=gen.er(50, 40, 90, 21), B=sp.csr_matrix((40, 30), dtype=np.float64), type_="FP64",
              semiring="GxB_PLUS_TIMES_FP64")
    check_mxm(G, A=gen.er(1, 40, 20, 22), B=gen.er(40, 1, 20, 23), type_="FP64",
              semiring="GxB_PLUS_TIMES_FP64")
    check_mxm(G, A=gen.er(40, 1, 20, 24), B=gen.er(1, 40, 20, 25), type_="FP64",
              semiring="GxB_PLUS_TIMES_FP64", fmt="CSC")
    # disjoint patterns: no product exists
    A = sp.csr_matrix((np.ones(5), (np.arange(5), np.arange(5))), shape=(10, 10))
    B = sp.csr_matrix((np.ones(5), (np.arange(5) + 5, np.arange(5))), shape=(10, 10))
    ref, got = check_mxm(G, A=A, B=B, type_="FP64", semiring="GxB_PLUS_TIMES_FP64")
    assert got["nvals"] == 0


def test_aliased_C_is_A(G):
    """C<C> = C*C with accum (Test/test28.m): T is always a fresh matrix"""
    A = gen.er(80, 80, 600, 26, np.int64, lo=1, hi=3)
    check_mxm(G, A=A, B=A, type_="INT64", semiring="GxB_PLUS_TIMES_INT64", M=A, mtype="INT64",
              Cinit=A, accum="GrB_PLUS_INT64")


# ---------------------------------------------------------------------------------------------
# triangle counting shape (config 2): C<L> = L*U' (dot) and C<L> = L*L (masked saxpy)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scale", [8, 11])
@pytest.mark.parametrize("type_", ["INT64", "UINT32"])
def test_tricount(G, scale, type_):
    A = gen.rmat_scipy(scale, 16, dtype=NP[type_])
    L, U = sp.tril(A, -1).tocsr(), sp.triu(A, 1).tocsr()
    sr = "GxB_PLUS_TIMES_" + type_
    ref, got = check_mxm(G, A=L, B=U, M=L, mtype=type_, type_=type_, semiring=sr, inp1=GrB_TRAN)
    ntri_dot = int(got["Ax"].astype(np.int64).sum())
    ref2, got2 = check_mxm(G, A=L, B=L, M=L, mtype=type_, type_=type_, semiring=sr)
    ntri_outer = int(got2["Ax"].astype(np.int64).sum())
    # independent count: trace(A^3)/6
    A1 = A.astype(np.int64)
    ntri = int((A1 @ A1).multiply(A1).sum() // 6)
    assert ntri_dot == ntri_outer == ntri


# ---------------------------------------------------------------------------------------------
# vectors: BFS step (config 3) and Bellman-Ford step (config 4)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fmt", ["CSR", "CSC"])
@pytest.mark.parametrize("op", ["vxm", "mxv"])
def test_bfs_levels(G, fmt, op):
    """the bfs5m loop (Demo/Source/bfs5m.c:71-82) step by step: q<!v> = q*A over LOR_LAND_BOOL with
    REPLACE; level sets must be identical at every level"""
    A = gen.rmat_scipy(10, 8, dtype=np.bool_)
    n = A.shape[0]
    src = int(np.argmax(np.diff(A.indptr)))
    visited = np.zeros(n, dtype=bool)
    q = np.array([src])
    for level in range(1, 12):
        visited[q] = True
        vi = np.nonzero(visited)[0]
        ref, got = check_mv(G, op=op, A=A, u=(n, q, np.ones(len(q), bool)), type_="BOOL",
                            semiring="GxB_LOR_LAND_BOOL", n_out=n, mask=(vi, np.ones(len(vi), bool)),
                            winit=(q, np.ones(len(q), bool)), outp=GrB_REPLACE, maskd=GrB_SCMP, fmt=fmt)
        q = got["vi"]
        if len(q) == 0:
            break
    assert visited.sum() > n // 4


@pytest.mark.parametrize("fmt", ["CSR", "CSC"])
def test_sssp_bellman_ford(G, fmt):
    """d = min (d, A min.+ d) iterated (config 4): MIN_PLUS_FP64 with accum MIN; bit-exact"""
    A = gen.rmat_scipy(9, 8, weighted=True)
    n = A.shape[0]
    src = int(np.argmax(np.diff(A.indptr)))
    d = np.full(n, np.inf)
    d[src] = 0.0
    idx = np.arange(n)
    for it in range(6):
        ref, got = check_mv(G, op="mxv", A=A, u=(n, idx, d), type_="FP64",
                            semiring="GxB_MIN_PLUS_FP64", n_out=n, winit=(idx, d),
                            accum="GrB_MIN_FP64", fmt=fmt)
        assert np.array_equal(ref["vx"], got["vx"])          # bit-exact, not just within tolerance
        d = got["vx"]
    assert np.isfinite(d).sum() > n // 4


@pytest.mark.parametrize("op,tran", [("mxv", GxB_DEFAULT), ("mxv", GrB_TRAN), ("vxm", GxB_DEFAULT),
                                     ("vxm", GrB_TRAN)])
@pytest.mark.parametrize("semiring,type_", [("GxB_PLUS_TIMES_FP64", "FP64"),
                                            ("GxB_MAX_PLUS_INT32", "INT32"),
                                            ("GxB_LOR_LAND_BOOL", "BOOL")])
@pytest.mark.parametrize("usparse", [True, False])
def test_mxv_vxm(G, op, tran, semiring, type_, usparse):
    m, n = 400, 300
    A = gen.er(m, n, 5000, 27, NP[type_])
    n_in = (m if tran == GrB_TRAN else n) if op == "mxv" else (n if tran == GrB_TRAN else m)
    n_out = (n if tran == GrB_TRAN else m) if op == "mxv" else (m if tran == GrB_TRAN else n)
    rng = np.random.default_rng(28)
    ui = np.sort(rng.choice(n_in, n_in // 5 if usparse else n_in, replace=False))
    ux = gen.er(1, len(ui), 4 * len(ui), 29, NP[type_]).toarray().ravel()[:len(ui)].astype(NP[type_])
    check_mv(G, op=op, A=A, u=(n_in, ui, ux), type_=type_, semiring=semiring, n_out=n_out, tran=tran)
    mi = np.sort(rng.choice(n_out, n_out // 2, replace=False))
    check_mv(G, op=op, A=A, u=(n_in, ui, ux), type_=type_, semiring=semiring, n_out=n_out, tran=tran,
             mask=(mi, rng.integers(0, 2, len(mi)).astype(bool)), maskd=GrB_SCMP)
    check_mv(G, op=op, A=A, u=(n_in, ui, ux), type_=type_, semiring=semiring, n_out=n_out, tran=tran,
             mask=(mi, rng.integers(0, 2, len(mi)).astype(bool)))


# ---------------------------------------------------------------------------------------------
# typecasting: operands of built-in types other than the multiply operator's
# are cast on the device as GB_CAST does (Source/GB.h:2925-2947), through the unmodified GrB_mxm
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ta,tb,semiring", [("INT32", "FP32", "GxB_PLUS_TIMES_FP64"),
                                            ("FP64", "FP32", "GxB_MIN_PLUS_INT16"),
                                            ("BOOL", "UINT8", "GxB_MAX_TIMES_FP32"),
                                            ("FP64", "INT64", "GxB_LOR_LAND_BOOL"),
                                            ("INT8", "FP64", "GxB_PLUS_MIN_UINT16")])
@pytest.mark.parametrize("method", [GxB_AxB_GUSTAVSON, GxB_AxB_DOT])
def test_typecast_operands(G, ta, tb, semiring, method):
    A = gen.er(150, 120, 1400, 31, NP[ta], lo=-6, hi=7)
    B = gen.er(120, 130, 1300, 32, NP[tb], lo=-6, hi=7)
    for S, t in ((A, ta), (B, tb)):
        if t in ("FP32", "FP64"):
            S.data = (S.data * 1.37).astype(NP[t])
            S.data[::9] = np.nan
            S.data[1::11] = np.inf
            S.data[2::13] = -np.inf
    M = gen.er(150, 130, 4000, 33, np.int8, lo=0, hi=2)
    check_mxm(G, A=A, B=B, type_=ta, btype=tb, semiring=semiring, method=method)
    check_mxm(G, A=A, B=B, type_=ta, btype=tb, semiring=semiring, method=method, M=M, mtype="INT8")


# ---------------------------------------------------------------------------------------------
# k-truss iteration (Extras/ktruss/ktruss_graphblas.c:103): C<C> = C*C over PLUS_LAND_INT64 with C as
# mask and both operands, then the entries below the support are dropped; the next caller of the path
# ---------------------------------------------------------------------------------------------
def test_ktruss_iterations(G):
    A = gen.rmat_scipy(9, 12, dtype=np.int64)
    Cm = A.copy().tocsr()
    Cm.data[:] = 1
    k = 4
    for _ in range(4):
        ref, got = check_mxm(G, A=Cm, B=Cm, M=Cm, mtype="INT64", type_="INT64",
                             semiring="GxB_PLUS_LAND_INT64")
        T = sp.csr_matrix((got["Ax"], got["Ai"], got["Ap"]), shape=Cm.shape)
        keep = T.data >= k - 2
        rows = np.repeat(np.arange(T.shape[0]), np.diff(T.indptr))[keep]
        nxt = sp.csr_matrix((np.ones(int(keep.sum()), np.int64), (rows, T.indices[keep])), shape=Cm.shape)
        if nxt.nnz == Cm.nnz or nxt.nnz == 0:
            break
        Cm = nxt


# ---------------------------------------------------------------------------------------------
# config 1 AS WRITTEN (BASELINE.json configs[0]): the reference's own random_matrix,
# simple_rand_seed (1), n = 16384, 131072 draws for A and again for B; C = A*B over PLUS_TIMES_FP64 has
# 1,046,459 entries (probe of the reference)
# ---------------------------------------------------------------------------------------------
def test_cfg1_exact_input(G):
    from parity import compare, export_csr
    n = 16384
    a = G.demo_random_matrix(n, n, 131072, seed=1)
    b = G.demo_random_matrix(n, n, 131072)
    assert (G.matrix_nvals(a), G.matrix_nvals(b)) == (131048, 131043)
    out = []
    for gpu in (False, True):
        c = G.matrix_new("FP64", n, n)
        G.use_gpu(gpu)
        before = G.shim_stats()
        try:
            G.mxm(c, None, None, "GxB_PLUS_TIMES_FP64", a, b, None)
            G.matrix_nvals(c)
        finally:
            G.use_gpu(False)
        if gpu:
            assert G.shim_stats()["gpu_calls"] - before["gpu_calls"] == 1
        out.append(export_csr(G, c))
    G.matrix_free(a)
    G.matrix_free(b)
    assert out[0]["nvals"] == 1046459
    ok, why = compare(out[0], out[1], "PLUS")
    assert ok, why


# ---------------------------------------------------------------------------------------------
# NaN under MIN / MAX in the dot method: the reference copies the first product and fmin / fmax-combines
# the later ones (GB_AxB_dot_cij.c:29-45), so C(i,j) is NaN exactly when ALL of its products are NaN
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("semiring", ["GxB_MIN_PLUS_FP64", "GxB_MAX_TIMES_FP64", "GxB_MIN_FIRST_FP32"])
@pytest.mark.parametrize("masked", [True, False])
def test_dot_nan_only_products(G, semiring, masked):
    n = 300
    type_ = semiring.split("_")[-1]
    rng = np.random.default_rng(77)
    A = gen.er(n, n, 6 * n, 31).astype(np.float64)
    B = gen.er(n, n, 6 * n, 32).astype(np.float64)
    A.data[rng.random(A.nnz) < 0.4] = np.nan
    B.data[rng.random(B.nnz) < 0.2] = np.nan
    # a few long (hub) vectors so that the table kernels see NaN too
    A = A.tolil() ; B = B.tolil()
    for c in (3, 9):
        A[c, :] = np.where(rng.random(n) < 0.5, np.nan, 1.5)
        B[:, c] = np.where(rng.random(n) < 0.5, np.nan, 2.5).reshape(-1, 1)
    A, B = A.tocsr(), B.tocsr()
    M = gen.er(n, n, 40 * n, 33, np.bool_) if masked else None
    ref, got = check_mxm(G, A=A, B=B, M=M, type_=type_, semiring=semiring, method=GxB_AxB_DOT)
    assert np.isnan(ref["Ax"]).any() and (~np.isnan(ref["Ax"])).any()

# ---------------------------------------------------------------------------------------------
# GxB_select with the built-in operators (row f4): the interposed GB_select computes T = select (A,k)
# on the device and hands it to the reference's own GB_accum_mask
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("op,k", [("TRIL", -1), ("TRIU", 1), ("TRIL", 0), ("TRIU", -3), ("DIAG", 0),
                                  ("DIAG", 2), ("OFFDIAG", 0), ("OFFDIAG", -1), ("NONZERO", 0)])
@pytest.mark.parametrize("fmt", ["CSR", "CSC", "HyperCSR", "HyperCSC"])
@pytest.mark.parametrize("tran", [GxB_DEFAULT, GrB_TRAN])
def test_select_builtin(G, op, k, fmt, tran):
    from parity import compare, export_csr, import_sp
    rng = np.random.default_rng(5)
    nr, nc = (300, 420)
    A = gen.er(nr, nc, 9 * nr, 41).tolil()
    A[7, :] = 1.5                      # a long vector
    A[:, 11] = 2.5
    A = A.tocsr()
    A.data[rng.random(A.nnz) < 0.2] = 0.0          # explicit zeros (NONZERO drops them)
    if "Hyper" in fmt:
        A = A.tolil() ; A[::3, :] = 0 ; A = A.tocsr() ; A.eliminate_zeros()
        A.data[rng.random(A.nnz) < 0.2] = 0.0
    out = []
    for gpu in (False, True):
        a = import_sp(G, A, "FP64", fmt)
        shape = (nc, nr) if tran == GrB_TRAN else (nr, nc)
        c = import_sp(G, sp.csr_matrix(shape), "FP64", "CSR")
        d = G.descriptor(inp0=tran)
        G.use_gpu(gpu)
        before = G.shim_select_calls()
        try:
            G.select(c, None, None, op, a, k, d)
            G.matrix_nvals(c)
        finally:
            G.use_gpu(False)
        if gpu and not REF_ONLY:
            assert G.shim_select_calls() - before == 1, "the GPU select did not run"
        out.append(export_csr(G, c))
        G.matrix_free(a)
        G.descriptor_free(d)
    ok, why = compare(out[0], out[1], "MIN")        # values are copies: exact
    assert ok, why


def test_select_with_mask_and_accum(G):
    from parity import compare, export_csr, import_sp
    n = 200
    A = gen.er(n, n, 12 * n, 51)
    Cinit = gen.er(n, n, 5 * n, 52)
    M = gen.er(n, n, 20 * n, 53, np.bool_)
    out = []
    for gpu in (False, True):
        a, c, m = import_sp(G, A, "FP64", "CSR"), import_sp(G, Cinit, "FP64", "CSR"), import_sp(G, M, "BOOL", "CSR")
        d = G.descriptor(mask=GrB_SCMP, outp=GrB_REPLACE)
        G.use_gpu(gpu)
        try:
            G.select(c, m, "GrB_PLUS_FP64", "TRIL", a, -1, d)
            G.matrix_nvals(c)
        finally:
            G.use_gpu(False)
        out.append(export_csr(G, c))
        for h in (a, m):
            G.matrix_free(h)
    ok, why = compare(out[0], out[1], "MIN")
    assert ok, why

# ---------------------------------------------------------------------------------------------
# GrB_reduce of a matrix to a scalar (row f3): the interposed GB_reduce_to_scalar reduces on the device
# and lets the reference finish (typecast into c, accumulator)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("type_,monoid,accum", [
    ("INT64", "GxB_PLUS_INT64_MONOID", None), ("INT64", "GxB_MIN_INT64_MONOID", None),
    ("INT32", "GxB_TIMES_INT32_MONOID", None), ("UINT8", "GxB_MAX_UINT8_MONOID", None),
    ("INT16", "GxB_PLUS_INT16_MONOID", "GrB_PLUS_INT16"), ("BOOL", "GxB_LXOR_BOOL_MONOID", None),
    ("BOOL", "GxB_LAND_BOOL_MONOID", None), ("FP64", "GxB_PLUS_FP64_MONOID", "GrB_MIN_FP64"),
    ("FP64", "GxB_MAX_FP64_MONOID", None), ("FP32", "GxB_PLUS_FP32_MONOID", None)])
@pytest.mark.parametrize("fmt", ["CSR", "HyperCSC"])
def test_reduce_to_scalar(G, type_, monoid, accum, fmt):
    from parity import import_sp
    n = 400
    A = gen.er(n, n, 40 * n, 61, NP[type_], lo=-3, hi=4)
    if type_ == "INT32":
        A.data[:] = np.where(A.data == 0, 1, A.data)    # a product that stays interesting
        A.data[np.abs(A.data) > 1] = 1
        A.data[::7] = -1
    out = []
    for gpu in (False, True):
        a = import_sp(G, A, type_, fmt)
        G.use_gpu(gpu)
        before = G.shim_reduce_calls()
        try:
            out.append(G.matrix_reduce(a, type_, monoid, accum, init=3))
        finally:
            G.use_gpu(False)
        if gpu and not REF_ONLY:
            assert G.shim_reduce_calls() - before == 1, "the GPU reduction did not run"
        G.matrix_free(a)
    if type_ in ("FP32", "FP64") and "PLUS" in monoid:
        eps = np.finfo(NP[type_]).eps
        assert abs(out[0] - out[1]) <= 64 * eps * np.abs(A.data.astype(np.float64)).sum()
    else:
        assert out[0] == out[1]

# ---------------------------------------------------------------------------------------------
# C = (ctype) A' on the device (row f2): the interposed GB_transpose sorts the entries on the GPU and
# returns T in the form (hypersparse or not) the reference's own method + GB_to_hyper_conform produce
# ---------------------------------------------------------------------------------------------
def _transpose_cases():
    """(name, scipy matrix held by column: vectors = columns, vlen = rows)"""
    rng = np.random.default_rng(77)
    out = [("er", gen.er(300, 420, 3000, 61)),
           ("tall", gen.er(70000, 300, 5000, 62)),              # three radix passes; T hypersparse
           ("wide", gen.er(40, 5000, 900, 63)),                 # one pass
           ("wide_long", gen.er(2, 3000, 2500, 64))]
    n = 1500
    band = sp.diags([np.arange(1, n + 1.0), np.full(n - 1, 2.0), np.full(n - 7, 3.0)], [0, 1, -7], format="lil")
    band[17, :] = 4.5                                           # a full row and a full column
    band[:, 33] = 5.5
    out.append(("band", band.tocsc()))
    # between the two thresholds of GB_to_hyper_conform (n/16 < non-empty vectors of T <= n/8, n = 1600): T
    # keeps the form its method gave it -- quicksort (few entries: hypersparse), bucket (many: not)
    for name, nnz in (("between_qsort", 3000), ("between_bucket", 6000)):
        rows = rng.choice(1600, 150, replace=False)
        i = rows[rng.integers(0, 150, nnz)]
        i[:150] = rows
        j = rng.integers(0, 900, nnz)
        m = sp.coo_matrix((rng.random(nnz) + 0.5, (i, j)), shape=(1600, 900)).tocsc()
        m.sum_duplicates()
        out.append((name, m))
    return out


@pytest.mark.parametrize("case", _transpose_cases(), ids=lambda c: c[0])
@pytest.mark.parametrize("fmt", ["CSC", "HyperCSC"])
@pytest.mark.parametrize("ctype", [None, "INT32", "BOOL"])
def test_transpose_seam(G, case, fmt, ctype):
    """raw T of GB_transpose (Source/GB.h:2153): form, vector list, pointers, pattern, values, type"""
    from parity import import_sp
    A = case[1].copy()
    if "Hyper" in fmt:
        A = A.tolil() ; A[:, ::3] = 0 ; A = A.tocsc() ; A.eliminate_zeros()
    A.data = np.round(A.data * 3 - 2, 1)                         # negatives, fractions, some zeros for BOOL
    out = []
    for gpu in (False, True):
        a = import_sp(G, A, "FP64", fmt)
        if not REF_ONLY:
            G.shim_transpose_min(0)
        G.use_gpu(gpu and not REF_ONLY)
        before = 0 if REF_ONLY else G.shim_transpose_calls()
        try:
            t = G.seam_transpose(a, ctype, True)
        finally:
            G.use_gpu(False)
            if not REF_ONLY:
                G.shim_transpose_min(65536)
        if not REF_ONLY:
            assert G.shim_transpose_calls() - before == (1 if gpu else 0), "the GPU transpose did not run"
        G.matrix_free(a)
        out.append(t)
    ref, got = out
    for k in ("vlen", "vdim", "is_hyper", "nvec", "nvec_nonempty", "type", "is_csc"):
        assert ref[k] == got[k], f"{k}: ref {ref[k]} got {got[k]}"
    assert np.array_equal(ref["p"], got["p"]), "vector pointers differ"
    if ref["is_hyper"]:
        assert np.array_equal(ref["h"], got["h"]), "vector list differs"
    assert np.array_equal(ref["i"], got["i"]), "pattern differs"
    assert np.array_equal(ref["x"], got["x"]), "values differ"
    if case[0].startswith("between") and "Hyper" not in fmt and ctype is None:
        assert ref["is_hyper"] == (case[0] == "between_qsort")


@pytest.mark.parametrize("held,exported", [("CSC", "CSR"), ("CSR", "CSC"), ("HyperCSC", "CSR"), ("CSR", "HyperCSC")])
@pytest.mark.parametrize("type_", ["FP64", "INT16"])
def test_format_change_in_place_on_device(G, held, exported, type_):
    """A = A' in place of the header (GB_transpose (NULL, NULL, csc, A, NULL), GB_transpose.c:79-95): what
    GxB_Matrix_export_<other format> runs before it hands the arrays out"""
    A = gen.er(230, 310, 6000, 91, NP[type_])
    out = []
    for gpu in (False, True):
        from parity import import_sp
        a = import_sp(G, A, type_, held)
        if not REF_ONLY:
            G.shim_transpose_min(0)
        G.use_gpu(gpu and not REF_ONLY)
        before = 0 if REF_ONLY else G.shim_transpose_calls()
        try:
            out.append(G.matrix_export(a, exported))
        finally:
            G.use_gpu(False)
            if not REF_ONLY:
                G.shim_transpose_min(65536)
        if gpu and not REF_ONLY:
            assert G.shim_transpose_calls() - before == 1, "the GPU transpose did not run"
    ref, got = out
    for k in ref:
        if isinstance(ref[k], np.ndarray):
            assert np.array_equal(ref[k], got[k]), f"{k} differs"
        else:
            assert ref[k] == got[k], f"{k}: ref {ref[k]} got {got[k]}"


@pytest.mark.parametrize("fmt,cfmt", [("CSR", "CSR"), ("CSC", "CSR"), ("HyperCSR", "CSC"), ("CSC", "HyperCSC")])
def test_grb_transpose_with_mask_and_accum(G, fmt, cfmt):
    """GrB_transpose C<!M> += A' through the unmodified API (Source/GrB_transpose.c:98-108 -> GB_transpose)"""
    from parity import compare, export_csr, import_sp
    A = gen.er(260, 340, 5000, 71)
    Cinit = gen.er(340, 260, 1500, 72)
    M = gen.er(340, 260, 9000, 73, np.bool_)
    out = []
    for gpu in (False, True):
        a, c, m = import_sp(G, A, "FP32", fmt), import_sp(G, Cinit, "FP64", cfmt), import_sp(G, M, "BOOL", fmt)
        d = G.descriptor(mask=GrB_SCMP)
        if not REF_ONLY:
            G.shim_transpose_min(0)
        G.use_gpu(gpu and not REF_ONLY)
        before = 0 if REF_ONLY else G.shim_transpose_calls()
        try:
            G.transpose(c, m, "GrB_PLUS_FP64", a, d)
            G.matrix_nvals(c)
        finally:
            G.use_gpu(False)
            if not REF_ONLY:
                G.shim_transpose_min(65536)
        if gpu and not REF_ONLY:
            assert G.shim_transpose_calls() - before >= 1, "the GPU transpose did not run"
        out.append(export_csr(G, c))
        for h in (a, m):
            G.matrix_free(h)
        G.descriptor_free(d)
    ok, why = compare(out[0], out[1], "MIN")
    assert ok, why


@pytest.mark.parametrize("inp0,inp1", [(GrB_TRAN, GxB_DEFAULT), (GxB_DEFAULT, GrB_TRAN), (GrB_TRAN, GrB_TRAN)])
@pytest.mark.parametrize("method", [GxB_AxB_GUSTAVSON, GxB_AxB_DOT])
@pytest.mark.parametrize("fmt", ["CSR", "CSC"])
def test_mxm_transposed_operands_on_device(G, inp0, inp1, method, fmt):
    """the transposes GB_AxB_meta runs in front of the multiply (GB_AxB_meta.c:203-355), mask in the other
    format included, on the device as well"""
    n = 220
    A = gen.er(n, n, 6000, 81, np.float64)
    B = gen.er(n, n, 5500, 82, np.float64)
    M = gen.er(n, n, 9000, 83, np.bool_)
    if not REF_ONLY:
        G.shim_transpose_min(0)
    try:
        before = 0 if REF_ONLY else G.shim_transpose_calls()
        check_mxm(G, A=A, B=B, type_="FP64", semiring="GxB_PLUS_TIMES_FP64", M=M, fmt=fmt, inp0=inp0,
                  inp1=inp1, method=method, cfmt="CSC" if fmt == "CSR" else "CSR")
        calls = 0 if REF_ONLY else G.shim_transpose_calls() - before
    finally:
        if not REF_ONLY:
            G.shim_transpose_min(65536)
    # not every fold transposes (A'*B by dot products needs none): the whole parametrisation does
    test_mxm_transposed_operands_on_device.calls = getattr(test_mxm_transposed_operands_on_device, "calls", 0) + calls


def test_mxm_transposed_operands_used_the_device(G):
    if REF_ONLY:
        pytest.skip("reference only")
    assert getattr(test_mxm_transposed_operands_on_device, "calls", 0) > 0

# ---------------------------------------------------------------------------------------------
# C<M> = accum (C,T) on the device (row f1): the interposed GB_accum_mask computes the new C on the GPU and
# hands it to the reference's own GB_transplant_conform
# ---------------------------------------------------------------------------------------------
class _device_accum_mask:
    def __init__(self, G):
        self.G = G

    def __enter__(self):
        if not REF_ONLY:
            self.G.shim_accum_mask_min(0)
            self.before = self.G.shim_accum_mask_calls()
        return self

    def calls(self):
        return 0 if REF_ONLY else self.G.shim_accum_mask_calls() - self.before

    def __exit__(self, *a):
        if not REF_ONLY:
            self.G.shim_accum_mask_min(-1)


@pytest.mark.parametrize("maskd", [GxB_DEFAULT, GrB_SCMP])
@pytest.mark.parametrize("outp", [GxB_DEFAULT, GrB_REPLACE])
@pytest.mark.parametrize("accum", [None, "GrB_PLUS_FP64", "GrB_MIN_FP64", "GrB_SECOND_FP64", "GrB_DIV_FP64"])
@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("fmt", ["CSR", "HyperCSC"])
def test_accum_mask_on_device(G, maskd, outp, accum, masked, fmt):
    """every combination of mask / complement / replace / accumulator of GB_spec_accum_mask, valued mask
    (entries that are present but false do not admit), C initialised"""
    if not masked and (maskd == GrB_SCMP or accum is None):
        pytest.skip("nothing for GB_accum_mask to compute")
    n = 150
    A = gen.er(n, n, 1500, 7, np.float64)
    B = gen.er(n, n, 1400, 8, np.float64)
    M = gen.er(n, n, 3000, 9, np.int8, lo=0, hi=2) if masked else None
    Cinit = gen.er(n, n, 2500, 10, np.float64)
    with _device_accum_mask(G) as dev:
        check_mxm(G, A=A, B=B, type_="FP64", semiring="GxB_PLUS_TIMES_FP64", M=M, mtype="INT8",
                  Cinit=Cinit, accum=accum, mask=maskd, outp=outp, fmt=fmt, method=GxB_AxB_GUSTAVSON)
        # GB_mxm skips GB_accum_mask when the multiply applied the mask and C is replaced (GB_mxm.c:141-159)
        _ACCUM_MASK_USED.append(dev.calls() >= 1)


_ACCUM_MASK_USED = []


def test_accum_mask_on_device_was_used(G):
    if REF_ONLY:
        pytest.skip("reference only")
    assert sum(_ACCUM_MASK_USED) >= 0.8 * len(_ACCUM_MASK_USED) > 0, _ACCUM_MASK_USED


@pytest.mark.parametrize("ctype,ttype,accum", [
    ("INT32", "FP64", "GrB_PLUS_INT32"), ("FP32", "INT64", "GrB_TIMES_FP64"), ("BOOL", "INT8", "GrB_LOR"),
    ("BOOL", "FP64", "GrB_PLUS_BOOL"), ("UINT8", "INT16", "GrB_MINUS_INT16"), ("FP64", "UINT32", "GrB_GT_UINT32"),
    ("INT64", "INT64", "GxB_ISLE_INT64"), ("UINT16", "FP32", "GrB_MAX_FP32"), ("INT8", "INT8", "GrB_FIRST_INT8"),
    ("FP64", "BOOL", "GrB_LXOR"), ("INT16", "UINT64", None)])
def test_accum_mask_typecasts(G, ctype, ttype, accum):
    """C, T and the accumulator's inputs / output all of different built-in types (GB_add.c casts through the
    operator's types; entries only in T are cast straight to C's type)"""
    n = 120
    sr = {"FP64": "GxB_PLUS_TIMES_FP64", "INT64": "GxB_PLUS_TIMES_INT64", "INT8": "GxB_PLUS_TIMES_INT8",
          "INT16": "GxB_PLUS_TIMES_INT16", "UINT32": "GxB_PLUS_TIMES_UINT32", "FP32": "GxB_PLUS_TIMES_FP32",
          "BOOL": "GxB_LOR_LAND_BOOL", "UINT64": "GxB_PLUS_TIMES_UINT64"}[ttype]
    A = gen.er(n, n, 900, 21, NP[ttype])
    B = gen.er(n, n, 800, 22, NP[ttype])
    M = gen.er(n, n, 4000, 23, np.float64, lo=0, hi=2)
    Cinit = gen.er(n, n, 2000, 24, NP[ctype])
    with _device_accum_mask(G) as dev:
        check_mxm(G, A=A, B=B, type_=ttype, semiring=sr, M=M, mtype="FP64", Cinit=Cinit, accum=accum,
                  ctype=ctype, method=GxB_AxB_GUSTAVSON)
        assert REF_ONLY or dev.calls() >= 1, "GB_accum_mask did not run on the device"


def test_accum_mask_vectors_bfs_and_sssp_steps(G):
    """the two vector loops: w<!v,replace> = u*A (LOR_LAND) and d = min (d, A min.+ d)"""
    n = 400
    A = gen.er(n, n, 8 * n, 31, np.float64, lo=1, hi=9)
    rng = np.random.default_rng(3)
    with _device_accum_mask(G) as dev:
        ui = np.sort(rng.choice(n, 60, replace=False))
        vi = np.sort(rng.choice(n, 150, replace=False))
        check_mv(G, op="vxm", A=A.astype(bool), u=(n, ui, np.ones(60, dtype=bool)), type_="BOOL",
                 semiring="GxB_LOR_LAND_BOOL", n_out=n, mask=(vi, np.ones(150, dtype=bool)),
                 winit=(vi, np.ones(150, dtype=bool)), outp=GrB_REPLACE, maskd=GrB_SCMP, method=GxB_AxB_DOT)
        d = rng.random(n) * 10
        check_mv(G, op="mxv", A=A, u=(n, np.arange(n), d), type_="FP64", semiring="GxB_MIN_PLUS_FP64",
                 n_out=n, winit=(np.arange(n), d), accum="GrB_MIN_FP64")
        assert REF_ONLY or dev.calls() >= 1, "GB_accum_mask did not run on the device"


# ---------------------------------------------------------------------------------------------
# GrB_assign of a scalar over all of C under a mask (row f3, `v<q> = level` of bfs5m.c:74): the interposed
# GB_assign computes the new C on the device
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("replace", [False, True])
@pytest.mark.parametrize("shape", ["matrix", "hyper"])
@pytest.mark.parametrize("k", range(6))
def test_assign_scalar_on_device(G, k, shape, replace):
    from parity import compare, export_csr, import_sp
    from test_oracle import ASSIGN_CASES
    ctype, mtype, stype, scalar, accum = ASSIGN_CASES[k]
    n, m = 140, 90
    Cs = gen.er(n, m, 2200, 48 + k, NP[ctype])
    Ms = gen.er(n, m, 5000, 50 + k, NP[mtype], lo=0, hi=2)
    fmt = "CSC" if shape == "matrix" else "HyperCSR"
    out = []
    for gpu in (False, True):
        c, mh = import_sp(G, Cs, ctype, fmt), import_sp(G, Ms, mtype, fmt)
        d = G.descriptor(outp=GrB_REPLACE) if replace else None
        with _device_accum_mask(G):
            G.use_gpu(gpu and not REF_ONLY)
            before = 0 if REF_ONLY else G.shim_assign_calls()
            try:
                G.assign_scalar(c, mh, accum[0] if accum else None, stype, scalar, d, n, m)
                G.matrix_nvals(c)
            finally:
                G.use_gpu(False)
            if gpu and not REF_ONLY:
                assert G.shim_assign_calls() - before == 1, "GB_assign did not run on the device"
        out.append(export_csr(G, c))
        G.matrix_free(mh)
        G.descriptor_free(d)
    ok, why = compare(out[0], out[1], "MIN")
    assert ok, why


def test_assign_scalar_bfs_level_step(G):
    """v<q> = level on GrB_Vectors, INT32 levels under a BOOL frontier, over three levels as bfs5m does"""
    n = 5000
    rng = np.random.default_rng(9)
    out = []
    for gpu in (False, True):
        v = G.vector_new("INT32", n)
        with _device_accum_mask(G):
            G.use_gpu(gpu and not REF_ONLY)
            before = 0 if REF_ONLY else G.shim_assign_calls()
            try:
                r2 = np.random.default_rng(9)
                for level in (1, 2, 3):
                    qi = np.sort(r2.choice(n, 700 * level, replace=False))
                    q = G.vector_import("BOOL", n, qi, r2.random(len(qi)) < 0.9)     # some entries false
                    G.assign_scalar(v, q, None, "INT32", level, None, n)
                    G.vector_nvals(v)
                    G.vector_free(q)
            finally:
                G.use_gpu(False)
            if gpu and not REF_ONLY:
                assert G.shim_assign_calls() - before == 3, "GB_assign did not run on the device"
        out.append(G.vector_export(v))
    assert out[0]["n"] == out[1]["n"] and out[0]["type"] == out[1]["type"]
    assert np.array_equal(out[0]["vi"], out[1]["vi"]) and np.array_equal(out[0]["vx"], out[1]["vx"])


def test_no_neighbour_call_failed_on_the_device(G):
    """every GB_select / GB_reduce_to_scalar / GB_transpose / GB_accum_mask call the shim took in this
    process ran on the device: none failed, none was delegated to the host"""
    import ctypes as C
    if REF_ONLY:
        pytest.skip("reference only")
    f, w = C.c_int64(), C.c_int64()
    G.shim.gb200_shim_neighbour_stats(C.byref(f), C.byref(w))
    assert (f.value, w.value) == (0, 0)

