"""GPU parity at the seam: libgb_b200.so called through its C ABI (gb200_AxB_host / gb200_AxB_device)
against the pinned oracle restatement on the same arrays, against the golden vectors generated from
the reference, and through size-independent properties at larger sizes."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import gen
import semirings
import graphblas_b200 as gb
import oracle_c

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NPT = {k: v[1] for k, v in gb.TYPES.items()}
FLOAT = ("FP32", "FP64")


def assert_same(ref: gb.Matrix, got: gb.Matrix, add: str, what=""):
    assert (ref.vlen, ref.vdim) == (got.vlen, got.vdim), what
    assert (ref.h is None) == (got.h is None), f"{what}: hypersparsity of T differs"
    assert np.array_equal(ref.p, got.p), f"{what}: vector pointers differ"
    if ref.h is not None:
        assert np.array_equal(ref.h, got.h), f"{what}: hyperlist differs"
    assert np.array_equal(ref.i, got.i), f"{what}: pattern differs"
    assert ref.type == got.type, what
    if ref.type in FLOAT and add in ("PLUS", "TIMES"):
        assert np.array_equal(np.isnan(ref.x), np.isnan(got.x)), f"{what}: NaN placement"
        inf = np.isinf(ref.x)
        assert np.array_equal(inf, np.isinf(got.x)) and np.array_equal(ref.x[inf], got.x[inf]), what
        fin = np.isfinite(ref.x)
        r, g = ref.x[fin].astype(np.float64), got.x[fin].astype(np.float64)
        eps = np.finfo(NPT[ref.type]).eps          # Test/GB_spec_compare.m: 64*eps, relative 1-norm
        assert np.abs(r - g).sum() <= 64 * eps * np.abs(r).sum(), f"{what}: beyond 64 eps"
    else:
        assert np.array_equal(ref.x, got.x, equal_nan=True), f"{what}: values differ (must be exact)"


def test_all_960_semirings_saxpy_and_dot():
    """every built-in worker on the GPU: saxpy (unmasked + masked) and dot (masked + unmasked),
    both flipxy settings for the operators that observe it"""
    n = 0
    for add, mult, t in semirings.all_builtin():
        dt = NPT[t]
        A = gb.Matrix.from_scipy(gen.er(40, 30, 260, 61, dt, lo=-3, hi=4).tocsc())
        B = gb.Matrix.from_scipy(gen.er(30, 35, 240, 62, dt, lo=-3, hi=4).tocsc())
        At = gb.Matrix.from_scipy(gen.er(30, 40, 260, 63, dt, lo=-3, hi=4).tocsc())
        M = gb.Matrix.from_scipy(gen.er(40, 35, 500, 64, np.bool_).tocsc())
        for flip in ((False, True) if mult in ("MINUS", "DIV", "FIRST", "SECOND", "GT", "ISGE") else (False,)):
            sr = gb.Semiring(add, mult, t, flip)
            tag = f"{add}_{mult}_{t} flip={flip}"
            assert_same(oracle_c.axb(None, False, A, B, sr), gb.axb_host(None, False, A, B, sr).matrix, add, tag + " saxpy")
            assert_same(oracle_c.axb(M, False, A, B, sr), gb.axb_host(M, False, A, B, sr).matrix, add, tag + " masked saxpy")
            assert_same(oracle_c.axb(M, False, At, B, sr, True), gb.axb_host(M, False, At, B, sr, True).matrix, add, tag + " masked dot")
            assert_same(oracle_c.axb(None, False, At, B, sr, True), gb.axb_host(None, False, At, B, sr, True).matrix, add, tag + " dot")
        n += 1
    assert n == 960


def test_golden_vectors_on_gpu():
    gdir = os.path.join(ROOT, "tests", "golden")
    files = sorted(f for f in os.listdir(gdir) if f.startswith("seam_") and f.endswith(".npz"))
    assert len(files) >= 10
    for f in files:
        z = np.load(os.path.join(gdir, f), allow_pickle=False)

        def mat(pfx):
            if pfx + "_p" not in z:
                return None
            h = z[pfx + "_h"] if pfx + "_h" in z else None
            return gb.Matrix(int(z[pfx + "_vlen"]), int(z[pfx + "_vdim"]), z[pfx + "_p"], z[pfx + "_i"],
                             z[pfx + "_x"], h, str(z[pfx + "_type"]))
        A, B, M, T = mat("A"), mat("B"), mat("M"), mat("T")
        sr = gb.Semiring(str(z["add"]), str(z["mult"]), str(z["xytype"]), bool(z["flipxy"]))
        r = gb.axb_host(M, bool(z["mask_comp"]), A, B, sr, bool(z["do_adotb"]))
        assert_same(T, r.matrix, sr.add, f)
        assert r.info["mask_applied"] == int(z["mask_applied"]), f
        assert (r.info["method_used"] == gb.METHOD_DOT) == bool(z["do_adotb"]), f


def test_tri_demo_known_answers_on_gpu():
    """triangle counts of the reference's Demo/Output/tri_demo.out (:66-1048) through both multiplies"""
    sr = gb.Semiring("PLUS", "TIMES", "INT64", True)
    gdir = os.path.join(ROOT, "tests", "golden")
    for f in sorted(os.listdir(gdir)):
        if not (f.startswith("tri_") and f.endswith(".npz")):
            continue
        z = np.load(os.path.join(gdir, f))
        n = int(z["n"])
        A = sp.csr_matrix((np.ones(len(z["i"]), np.int64), z["i"], z["p"]), shape=(n, n))
        L, U = gb.Matrix.from_scipy(sp.tril(A, -1).tocsr()), gb.Matrix.from_scipy(sp.triu(A, 1).tocsr())
        dot = gb.axb_host(L, False, U, L, sr, True).matrix
        outer = gb.axb_host(L, False, L, L, sr, False).matrix
        assert int(dot.x.sum()) == int(outer.x.sum()) == int(z["ntri"]), f


@pytest.mark.parametrize("hyper", [False, True])
@pytest.mark.parametrize("masked", [False, True])
def test_flopcount_matches_oracle(hyper, masked):
    A = gb.Matrix.from_scipy(gen.er(3000, 2000, 30000, 71).tocsc())
    B = gb.Matrix.from_scipy(gen.er(2000, 2500, 28000, 72).tocsc())
    B2 = gen.er(2000, 2500, 28000, 72).tolil()
    B2[:, 7] = 1.0                                  # one vector longer than the multi-block threshold
    B2[::2, 9] = 1.0
    B = gb.Matrix.from_scipy(B2.tocsc())
    M = gb.Matrix.from_scipy(gen.er(3000, 2500, 9000, 73, np.bool_).tocsc())
    if hyper:
        A, B, M = A.to_hyper(), B.to_hyper(), M.to_hyper()
    dA, dB, dM = gb.DMatrix(A), gb.DMatrix(B), gb.DMatrix(M)
    got, total = gb.flopcount(dM if masked else None, dA, dB)
    ref, rtotal = oracle_c.flopcount(M if masked else None, A, B)
    assert total == rtotal and np.array_equal(got, ref)


@pytest.mark.parametrize("scale", [14, 16])
def test_rmat_spgemm_and_tricount_properties(scale):
    """larger RMAT inputs: (1) full parity of C<L>=L*U' against the oracle, (2) dot == masked saxpy
    == trace(A^3)/6, (3) unmasked C=A*A on a thinned graph against the oracle (heavy bitmap bins)"""
    A = gen.rmat_scipy(scale, 8, dtype=np.int64)
    L, U = gb.Matrix.from_scipy(sp.tril(A, -1).tocsr()), gb.Matrix.from_scipy(sp.triu(A, 1).tocsr())
    sr = gb.Semiring("PLUS", "TIMES", "INT64", True)
    dot = gb.axb_host(L, False, U, L, sr, True)
    outer = gb.axb_host(L, False, L, L, sr, False)
    assert_same(oracle_c.axb(L, False, U, L, sr, True), dot.matrix, "PLUS", "tri dot")
    assert_same(oracle_c.axb(L, False, L, L, sr, False), outer.matrix, "PLUS", "tri outer")
    assert int(dot.matrix.x.sum()) == int(outer.matrix.x.sum())
    assert dot.info["flops"] == int(dot.matrix.x.sum())         # matched pairs == triangles
    if scale <= 14:
        Am = gb.Matrix.from_scipy(A.tocsr())
        got = gb.axb_host(None, False, Am, Am, gb.Semiring("PLUS", "TIMES", "INT64"))
        assert_same(oracle_c.axb(None, False, Am, Am, gb.Semiring("PLUS", "TIMES", "INT64")), got.matrix,
                    "PLUS", "A*A")
        fl, total = oracle_c.flopcount(None, Am, Am)
        assert got.info["flops"] == total


def test_resident_operands_repeatable():
    """gb200_AxB_device on resident operands: integer results identical run to run and identical to
    the host entry point; floating PLUS within tolerance"""
    A = gb.Matrix.from_scipy(gen.er(5000, 5000, 60000, 81, np.int64).tocsc())
    dA = gb.DMatrix(A)
    sr = gb.Semiring("PLUS", "TIMES", "INT64")
    r1 = gb.axb_device(None, False, dA, dA, sr).matrix
    r2 = gb.axb_device(None, False, dA, dA, sr).matrix
    r3 = gb.axb_host(None, False, A, A, sr).matrix
    for r in (r2, r3):
        assert np.array_equal(r1.p, r.p) and np.array_equal(r1.i, r.i) and np.array_equal(r1.x, r.x)


def test_mask_value_types():
    """a mask may have any built-in type; an entry is true iff its value is nonzero (NaN is true)"""
    A = gb.Matrix.from_scipy(gen.er(200, 150, 2500, 91).tocsc())
    B = gb.Matrix.from_scipy(gen.er(150, 180, 2200, 92).tocsc())
    sr = gb.Semiring("PLUS", "TIMES", "FP64")
    for t in ("BOOL", "INT8", "UINT16", "INT32", "UINT64", "FP32", "FP64"):
        Ms = gen.er(200, 180, 9000, 93, NPT[t], lo=0, hi=2)
        if t in FLOAT:
            Ms.data[::7] = np.nan
            Ms.data[1::7] = -0.0
        M = gb.Matrix.from_scipy(Ms.tocsc(), t)
        for comp in (False, True):
            for dot, Ax in ((False, A), (True, gb.Matrix.from_scipy(gen.er(150, 200, 2500, 94).tocsc()))):
                assert_same(oracle_c.axb(M, comp, Ax, B, sr, dot), gb.axb_host(M, comp, Ax, B, sr, dot).matrix,
                            "PLUS", f"mask {t} comp={comp} dot={dot}")


def test_nan_inf_placement():
    """NaN / Inf must land in the same places as in the reference (Test/isequal_roundoff.m:18-35)"""
    A = gen.er(120, 100, 1500, 95)
    B = gen.er(100, 110, 1400, 96)
    A.data[::11] = np.inf
    A.data[3::17] = -np.inf
    B.data[::13] = np.nan
    B.data[5::19] = 0.0
    Am, Bm = gb.Matrix.from_scipy(A.tocsc()), gb.Matrix.from_scipy(B.tocsc())
    for add, mult in (("PLUS", "TIMES"), ("MIN", "PLUS"), ("MAX", "TIMES"), ("TIMES", "MIN"), ("MIN", "DIV")):
        sr = gb.Semiring(add, mult, "FP64")
        assert_same(oracle_c.axb(None, False, Am, Bm, sr), gb.axb_host(None, False, Am, Bm, sr).matrix, add,
                    f"{add}_{mult}")


# ---------------------------------------------------------------------------------------------
# vector multiplies (B is n-by-1): the GrB_mxv / GrB_vxm shapes, pull (dot) and push (saxpy)
# ---------------------------------------------------------------------------------------------
def _vector(n, idx, x, t):
    return gb.Matrix(n, 1, np.array([0, len(idx)]), np.asarray(idx, np.int64), np.asarray(x, NPT[t]), None, t)


def _filter_by_mask(T: gb.Matrix, M: gb.Matrix, comp: bool) -> gb.Matrix:
    """what GB_mask leaves of an n-by-1 T under <M> or <!M> (valued mask)"""
    true_idx = M.i[M.x.astype(bool) | (np.isnan(M.x) if M.x.dtype.kind == "f" else False)]
    keep = np.isin(T.i, true_idx)
    if comp:
        keep = ~keep
    return gb.Matrix(T.vlen, 1, np.array([0, int(keep.sum())]), T.i[keep], T.x[keep], None, T.type)


@pytest.mark.parametrize("add,mult,t", [("PLUS", "TIMES", "FP64"), ("MIN", "PLUS", "FP64"),
                                        ("LOR", "LAND", "BOOL"), ("MAX", "MIN", "INT32"),
                                        ("PLUS", "TIMES", "UINT8"), ("LXOR", "GT", "INT16"),
                                        ("TIMES", "FIRST", "INT64"), ("MIN", "SECOND", "FP32")])
@pytest.mark.parametrize("dense_u", [False, True])
def test_vector_pull_and_push(add, mult, t, dense_u):
    """w = A'*u (dot) and w = A*u (saxpy) with no mask, M and !M, including vectors of A longer than
    the per-group limit (hub rows) so that the segment kernels run"""
    n = 6000
    rng = np.random.default_rng(97)
    As = gen.er(n, n, 10 * n, 98, NPT[t], lo=-3, hi=4).tolil()
    for r in (3, 777):                                       # two hub vectors (> 2048 entries)
        cols = rng.choice(n, 3000, replace=False)
        As[r, cols] = 1
        As[cols, r] = 1
    A = gb.Matrix.from_scipy(As.tocsc(), t)
    nu = n if dense_u else n // 7
    ui = np.sort(rng.choice(n, nu, replace=False))
    ux = rng.integers(-3, 4, nu) if t not in FLOAT else rng.random(nu) * 2 - 0.5
    if t == "BOOL":
        ux = rng.integers(0, 2, nu)
    if t.startswith("U"):
        ux = np.abs(ux)
    u = _vector(n, ui, ux, t)
    mi = np.sort(rng.choice(n, n // 3, replace=False))
    M = _vector(n, mi, rng.integers(0, 2, len(mi)), "INT8")
    sr = gb.Semiring(add, mult, t)
    for dot in (True, False):
        for mask, comp in ((None, False), (M, False), (M, True)):
            ref_info = {}
            ref = oracle_c.axb(mask, comp, A, u, sr, dot, ref_info)
            got = gb.axb_host(mask, comp, A, u, sr, dot)
            what = f"{add}_{mult}_{t} dot={dot} mask={'none' if mask is None else ('!M' if comp else 'M')}"
            if mask is not None and got.info["mask_applied"] and not ref_info["mask_applied"]:
                # the push kernel applies !M itself; the reference filters T afterwards (same final C)
                assert comp and not dot, what
                ref = _filter_by_mask(ref, mask, comp)
            else:
                assert bool(got.info["mask_applied"]) == bool(ref_info["mask_applied"]), what
            assert_same(ref, got.matrix, add, what)


def test_vector_push_unfused_compmask(monkeypatch):
    """GB200_FUSE_COMPMASK=0: the complemented mask is dropped exactly as the reference's saxpy does"""
    monkeypatch.setenv("GB200_FUSE_COMPMASK", "0")
    n = 3000
    A = gb.Matrix.from_scipy(gen.er(n, n, 8 * n, 99, np.bool_).tocsc(), "BOOL")
    rng = np.random.default_rng(100)
    ui = np.sort(rng.choice(n, 200, replace=False))
    u = _vector(n, ui, np.ones(200), "BOOL")
    mi = np.sort(rng.choice(n, 1000, replace=False))
    M = _vector(n, mi, np.ones(1000), "BOOL")
    sr = gb.Semiring("LOR", "LAND", "BOOL", True)
    info = {}
    ref = oracle_c.axb(M, True, A, u, sr, False, info)
    got = gb.axb_host(M, True, A, u, sr, False)
    assert got.info["mask_applied"] == 0 == info["mask_applied"]
    assert_same(ref, got.matrix, "LOR", "unfused !M")


# ---------------------------------------------------------------------------------------------
# masked dot with hub vectors: owners longer than one shared-memory table load (segments with
# cursors), a dense owner, walks longer than one task (split pairs), pattern-only (iso) operands
# and valued ones, pairs with short owners (the lane-group kernel)
# ---------------------------------------------------------------------------------------------
def _hub_matrix(n, seed, dtype, iso):
    rng = np.random.default_rng(seed)
    S = gen.er(n, n, 6 * n, seed, dtype, lo=1, hi=5).tolil()
    for c, ln in ((5, 9000), (17, 5000), (40, n), (41, 2500), (90, 1500)):
        rows = np.sort(rng.choice(n, ln, replace=False))
        S[rows, c] = rng.integers(1, 5, ln)
    S = S.tocsc()
    if iso:
        S.data[:] = 1
    return S


@pytest.mark.parametrize("iso", [True, False])
@pytest.mark.parametrize("add,mult,t", [("PLUS", "TIMES", "INT64"), ("MIN", "PLUS", "FP64"),
                                        ("LOR", "LAND", "BOOL"), ("TIMES", "PLUS", "INT32"),
                                        ("LXOR", "LOR", "BOOL"), ("PLUS", "SECOND", "FP32"),
                                        ("MAX", "MINUS", "INT8"), ("EQ", "GE", "UINT16")])
def test_masked_dot_hubs(iso, add, mult, t):
    n = 12000
    dt = NPT[t]
    A = gb.Matrix.from_scipy(_hub_matrix(n, 201, dt, iso), t)
    B = gb.Matrix.from_scipy(_hub_matrix(n, 202, dt, iso), t)
    rng = np.random.default_rng(203)
    Ms = gen.er(n, n, 4 * n, 204, np.bool_).tolil()
    hubs = [5, 17, 40, 41, 90]
    for i in hubs:                      # hub-by-hub pairs and hub-by-anything rows/columns
        for j in hubs:
            Ms[i, j] = True
        Ms[i, rng.choice(n, 300, replace=False)] = True
        Ms[rng.choice(n, 300, replace=False), i] = True
    M = gb.Matrix.from_scipy(Ms.tocsc(), "BOOL")
    sr = gb.Semiring(add, mult, t)
    for m in (M, M.to_hyper()):
        ref = oracle_c.axb(m, False, A, B, sr, True)
        got = gb.axb_host(m, False, A, B, sr, True)
        assert_same(ref, got.matrix, add, f"{add}_{mult}_{t} iso={iso}")
        assert got.info["mask_applied"] == 1 and got.info["method_used"] == gb.METHOD_DOT


@pytest.mark.parametrize("iso", [True, False])
@pytest.mark.parametrize("bits", ["4096", "1024", "512", "old"])
def test_masked_dot_hub_variants(monkeypatch, iso, bits):
    """the hub owners through (a) several bitmap parts of the row-walk kernel (GB200_DOTR_BM_BITS=4096:
    the 12000-wide index range takes 3 parts; 1024: 12 parts), (b) more parts than it accepts (512: the
    segmented cuckoo kernel serves them), (c) the round-1 kernels alone (GB200_DOTR=0): same T as the oracle"""
    if bits == "old":
        monkeypatch.setenv("GB200_DOTR", "0")
    else:
        monkeypatch.setenv("GB200_DOTR_BM_BITS", bits)
    for add, mult, t in (("PLUS", "TIMES", "INT64"), ("MIN", "PLUS", "FP64"), ("LXOR", "LOR", "BOOL")):
        test_masked_dot_hubs(iso, add, mult, t)


@pytest.mark.parametrize("iso", [True, False])
@pytest.mark.parametrize("env", [{"GB200_DOTG_TRIM": "0"}, {"GB200_DOTG_TRIM": "1"}, {"GB200_DOTG_TRIM": "2"},
                                 {"GB200_DOT_STREAMS": "0"}, {"GB200_DOT_STREAMS": "1", "GB200_DOTR_BM_BITS": "4096"}])
def test_masked_dot_setup_and_launch_variants(monkeypatch, iso, env):
    """the trim modes of the classification (none / always searched / not searched in a walked list of one
    row), the semiring kernels on side streams or one after another: every variant returns the oracle's T"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for add, mult, t in (("PLUS", "TIMES", "INT64"), ("MIN", "PLUS", "FP64")):
        test_masked_dot_hubs(iso, add, mult, t)
    # longer walked lists on both sides (the searches of the trim run on real ranges)
    A = gen.rmat_scipy(13, 16, dtype=np.int64)
    if not iso:
        A.data[:] = np.random.default_rng(5).integers(1, 9, A.nnz)
    L, U = gb.Matrix.from_scipy(sp.tril(A, -1).tocsr()), gb.Matrix.from_scipy(sp.triu(A, 1).tocsr())
    sr = gb.Semiring("PLUS", "TIMES", "INT64", True)
    ref = oracle_c.axb(L, False, U, L, sr, True)
    got = gb.axb_host(L, False, U, L, sr, True)
    assert_same(ref, got.matrix, "PLUS", f"tri {env}")


def test_masked_dot_general_path_on_iso_input(monkeypatch):
    """GB200_DOTG_ISO=0 sends pattern-only operands through the valued kernel: same T"""
    A = gen.rmat_scipy(13, 8, dtype=np.int64)
    L, U = gb.Matrix.from_scipy(sp.tril(A, -1).tocsr()), gb.Matrix.from_scipy(sp.triu(A, 1).tocsr())
    sr = gb.Semiring("PLUS", "TIMES", "INT64", True)
    fast = gb.axb_host(L, False, U, L, sr, True)
    monkeypatch.setenv("GB200_DOTG_ISO", "0")
    slow = gb.axb_host(L, False, U, L, sr, True)
    assert_same(fast.matrix, slow.matrix, "PLUS", "iso vs general")
    assert fast.info["flops"] == slow.info["flops"] == int(fast.matrix.x.sum())


# ---------------------------------------------------------------------------------------------
# typecasting at the seam: the library casts operands of other built-in types on the device
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ta,tb,add,mult,txy", [("INT32", "FP32", "PLUS", "TIMES", "FP64"),
                                                ("FP64", "FP64", "PLUS", "TIMES", "INT16"),
                                                ("FP32", "INT8", "MIN", "PLUS", "INT64"),
                                                ("BOOL", "UINT8", "MAX", "TIMES", "FP32"),
                                                ("FP64", "INT64", "LOR", "LAND", "BOOL"),
                                                ("UINT64", "INT8", "PLUS", "MIN", "UINT16"),
                                                ("FP32", "FP64", "LXOR", "GT", "INT32")])
def test_typecast_at_the_seam(ta, tb, add, mult, txy):
    def operand(r, c, nnz, seed, t):
        S = gen.er(r, c, nnz, seed, NPT[t], lo=-6, hi=7).tocsc()
        if t in FLOAT:
            S.data = (S.data * 1.37).astype(NPT[t])
            S.data[::9] = np.nan
            S.data[1::11] = np.inf
            S.data[2::13] = -np.inf
        return gb.Matrix.from_scipy(S, t)
    A, At, B = operand(400, 300, 5000, 401, ta), operand(300, 400, 5000, 402, ta), operand(300, 350, 4500, 403, tb)
    M = gb.Matrix.from_scipy(gen.er(400, 350, 9000, 404, np.int8, lo=0, hi=2).tocsc())
    sr = gb.Semiring(add, mult, txy)
    for mask in (None, M):
        assert_same(oracle_c.axb(mask, False, A, B, sr), gb.axb_host(mask, False, A, B, sr).matrix, add, "saxpy")
        assert_same(oracle_c.axb(mask, False, At, B, sr, True), gb.axb_host(mask, False, At, B, sr, True).matrix, add, "dot")


def test_masked_dot_edge_cases():
    """empty mask, a mask whose entries are all false, empty operands, a single pair, 1-by-n shapes"""
    n = 500
    A = gb.Matrix.from_scipy(gen.er(n, n, 9 * n, 501, np.int32, lo=1, hi=5).tocsc())
    B = gb.Matrix.from_scipy(gen.er(n, n, 9 * n, 502, np.int32, lo=1, hi=5).tocsc())
    sr = gb.Semiring("PLUS", "TIMES", "INT32")
    empty = gb.Matrix.from_scipy(sp.csc_matrix((n, n), dtype=np.bool_))
    Mf = gen.er(n, n, 6 * n, 503, np.int8, lo=0, hi=1).tocsc()     # present, all false
    Mf.data[:] = 0
    one = gb.Matrix.from_scipy(sp.csc_matrix((np.array([True]), (np.array([7]), np.array([9]))), shape=(n, n)))
    Ae = gb.Matrix.from_scipy(sp.csc_matrix((n, n), dtype=np.int32))
    M = gb.Matrix.from_scipy(gen.er(n, n, 6 * n, 504, np.bool_).tocsc())
    for m, a, b, what in ((empty, A, B, "empty mask"), (gb.Matrix.from_scipy(Mf, "INT8"), A, B, "all-false mask"),
                          (one, A, B, "one pair"), (M, Ae, B, "empty A"), (M, A, Ae, "empty B"),
                          (M.to_hyper(), A.to_hyper(), B.to_hyper(), "all hypersparse")):
        for comp in (False, True):
            if comp and what in ("empty mask", "all-false mask"):
                continue                                    # C<!empty> = A'*B is the full n*n dot product
            assert_same(oracle_c.axb(m, comp, a, b, sr, True), gb.axb_host(m, comp, a, b, sr, True).matrix,
                        "PLUS", f"{what} comp={comp}")
    row = gb.Matrix.from_scipy(gen.er(n, 1, 40, 505, np.int32, lo=1, hi=5).tocsc())
    col = gb.Matrix.from_scipy(gen.er(n, 3, 90, 506, np.int32, lo=1, hi=5).tocsc())
    m13 = gb.Matrix.from_scipy(sp.csc_matrix(np.ones((1, 3), dtype=np.bool_)))
    assert_same(oracle_c.axb(m13, False, row, col, sr, True), gb.axb_host(m13, False, row, col, sr, True).matrix,
                "PLUS", "1-by-3 result")


# ---------------------------------------------------------------------------------------------
# C = (ctype) A' through the C ABI (gb200_transpose_host, row f2) against the pinned restatement
# ---------------------------------------------------------------------------------------------
def _start_form(A: gb.Matrix, ctype) -> bool:
    """the form the reference's method leaves T in before GB_to_hyper_conform (GB_transpose.c:482-606)"""
    if A.h is not None:
        return True
    anz, csize = int(A.p[-1]), np.dtype(NPT[ctype or A.type]).itemsize
    return max(24.0 * anz, (16.0 + csize) * anz) < 16.0 * A.vlen + (8.0 + csize) * anz


@pytest.mark.parametrize("hyper", [False, True])
@pytest.mark.parametrize("ctype", [None, "INT16", "BOOL", "FP32", "UINT64"])
def test_transpose_matches_oracle(hyper, ctype):
    from test_oracle import transpose_inputs
    for name, S in transpose_inputs():
        S = S.copy()
        S.data = np.round(S.data * 300 - 150, 1)
        S.data[::7] = np.inf
        S.data[3::11] = np.nan
        A = gb.Matrix.from_scipy(S.tocsc(), "FP64")
        if hyper:
            A = A.to_hyper()
        ref = oracle_c.transpose(A, ctype)
        got = gb.transpose_host(A, ctype, hyper=_start_form(A, ctype), hyper_ratio=0.0625)
        assert_same(ref, got.matrix, "MIN", f"transpose {name}")
        assert got.info["nvec_nonempty"] == int(np.count_nonzero(np.diff(ref.p)))


def test_transpose_properties_rmat():
    """at a size the oracle does not run at: (A')' == A bit for bit, and A' == A for the symmetric graph"""
    A = gb.Matrix.from_scipy(gen.rmat_scipy(16, 16, weighted=True).tocsc(), "FP64")
    L = gb.select_host(A, "TRIL", -1).matrix                        # not symmetric
    T = gb.transpose_host(L, None, hyper=False).matrix
    TT = gb.transpose_host(T, None, hyper=False).matrix
    assert_same(L, TT, "MIN", "(L')'")
    U = gb.select_host(A, "TRIU", 1).matrix
    assert np.array_equal(T.p, U.p) and np.array_equal(T.i, U.i), "tril (A)' != triu (A) of a symmetric A"


# ---------------------------------------------------------------------------------------------
# C<M> = accum (C,T) through the C ABI (gb200_accum_mask_host, row f1) against the pinned restatement
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("comp", [False, True])
@pytest.mark.parametrize("replace", [False, True])
@pytest.mark.parametrize("hyper", [False, True])
def test_accum_mask_matches_oracle(comp, replace, hyper):
    from test_oracle import ACCUM_MASK_CASES, accum_mask_inputs
    for k, (ctype, ttype, mtype, accum) in enumerate(ACCUM_MASK_CASES):
        if mtype is None and comp:
            continue
        Cm, T, M = accum_mask_inputs(ctype, ttype, mtype, hyper, seed=k)
        acc = (accum[1], accum[2]) if accum else None
        ref = oracle_c.accum_mask(Cm, T, M, comp, replace, acc, hyper)
        got = gb.accum_mask_host(Cm, T, M, comp, replace, acc, hyper).matrix
        assert_same(ref, got, "MIN", f"accum_mask case {k} {ctype} {ttype} {mtype} {accum}")


def test_accum_mask_edge_cases():
    """empty C, empty T, empty mask, C aliased with the mask, a vector (vdim == 1), long vectors"""
    n = 300
    E = gb.Matrix.from_scipy(sp.csc_matrix((n, n)), "FP64")
    A = gb.Matrix.from_scipy(gen.er(n, n, 4000, 91).tocsc(), "FP64")
    B = gb.Matrix.from_scipy(gen.er(n, n, 3000, 92).tocsc(), "FP64")
    acc = ("PLUS", "FP64")
    for Cm, T, M, comp, rep, a in [(E, A, None, False, False, acc), (A, E, None, False, False, acc),
                                   (A, B, E, False, False, None), (A, B, E, True, True, acc),
                                   (A, B, A, False, True, None), (A, B, A, True, False, acc),
                                   (E, E, E, True, False, acc)]:
        ref = oracle_c.accum_mask(Cm, T, M, comp, rep, a, False)
        got = gb.accum_mask_host(Cm, T, M, comp, rep, a, False).matrix
        assert_same(ref, got, "MIN", "accum_mask edge case")
    # n-by-1 vectors: d = min (d, t) with all entries present, and a sparse frontier under a complemented mask
    rng = np.random.default_rng(5)
    nv = 50000
    def vec(idx, x, t):
        return gb.Matrix(nv, 1, np.array([0, len(idx)]), np.asarray(idx, dtype=np.int64), np.asarray(x), None, t)
    d = vec(np.arange(nv), rng.random(nv), "FP64")
    t = vec(np.sort(rng.choice(nv, 20000, replace=False)), rng.random(20000), "FP64")
    v = vec(np.sort(rng.choice(nv, 30000, replace=False)), np.ones(30000, dtype=np.bool_), "BOOL")
    for Cm, T, M, comp, rep, a in [(d, t, None, False, False, ("MIN", "FP64")), (t, d, v, True, True, None),
                                   (t, t, v, True, False, ("PLUS", "FP64"))]:
        ref = oracle_c.accum_mask(Cm, T, M, comp, rep, a, False)
        got = gb.accum_mask_host(Cm, T, M, comp, rep, a, False).matrix
        assert_same(ref, got, "MIN", "accum_mask on vectors")


@pytest.mark.parametrize("replace", [False, True])
@pytest.mark.parametrize("shape", ["matrix", "hyper", "vector"])
def test_assign_scalar_matches_oracle(replace, shape):
    """gb200_assign_scalar_host (row f3) against the pinned restatement: GB_accum_mask with T = the scalar on
    the pattern of the mask's true entries (pinned against GrB_*_assign in tests/test_oracle.py)"""
    from test_oracle import ASSIGN_CASES, accum_mask_inputs, scalar_on_mask
    for k, (ctype, mtype, stype, scalar, accum) in enumerate(ASSIGN_CASES):
        if shape == "vector":
            Cm = gb.Matrix.from_scipy(gen.er(3000, 1, 900, 61 + k, NPT[ctype]).tocsc(), ctype)
            M = gb.Matrix.from_scipy(gen.er(3000, 1, 1500, 62 + k, NPT[mtype], lo=0, hi=2).tocsc(), mtype)
        else:
            Cm, _, M = accum_mask_inputs(ctype, ctype, mtype, shape == "hyper", seed=k)
        acc = (accum[1], accum[2]) if accum else None
        ref = oracle_c.accum_mask(Cm, scalar_on_mask(M, scalar, stype), M, False, replace, acc, shape == "hyper")
        got = gb.assign_scalar_host(Cm, M, scalar, stype, replace, acc, shape == "hyper").matrix
        assert_same(ref, got, "MIN", f"assign case {k} {shape}")
