"""parity.py -- run one GrB_mxm / GrB_mxv / GrB_vxm twice through the reference's public API: once on
the reference's own CPU path (shim off) and once on the B200 (shim on), and compare the exported
results with the reference's own test criterion (Test/GB_spec_compare.m:18-24,40-52 and
Test/isequal_roundoff.m:18-47): identical pattern and type; integer / boolean values exact;
floating point within 64*eps relative 1-norm with identical NaN / Inf placement."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

import grbref
from grbref import (GxB_DEFAULT, GrB_REPLACE, GrB_SCMP, GrB_TRAN)

FLOAT_TYPES = ("FP32", "FP64")
# harness self-check on a box without a GPU: run both sides on the reference
REF_ONLY = bool(__import__("os").environ.get("GB200_TEST_REF_ONLY"))


def import_sp(G, m: sp.spmatrix, type_: str, fmt: str):
    """scipy matrix -> GrB_Matrix stored as CSR / CSC / HyperCSR / HyperCSC"""
    nrows, ncols = m.shape
    by_row = fmt in ("CSR", "HyperCSR")
    s = m.tocsr() if by_row else m.tocsc()
    s.sort_indices()
    x = s.data.astype(grbref.NP_OF[type_])
    if fmt in ("CSR", "CSC"):
        return G.matrix_import(fmt, type_, nrows, ncols, s.indptr, s.indices, x)
    cnt = np.diff(s.indptr)
    h = np.nonzero(cnt > 0)[0]
    p = np.concatenate([[0], np.cumsum(cnt[h])])
    return G.matrix_import(fmt, type_, nrows, ncols, p, s.indices, x, h)


def export_csr(G, A):
    return G.matrix_export(A, "CSR")


def values_equal(type_: str, add: str, ref: np.ndarray, got: np.ndarray) -> tuple[bool, str]:
    if ref.shape != got.shape:
        return False, f"value count {got.shape} != {ref.shape}"
    if type_ not in FLOAT_TYPES:
        bad = np.nonzero(ref != got)[0]
        if len(bad):
            return False, f"{len(bad)} values differ, first at {bad[0]}: ref {ref[bad[0]]} got {got[bad[0]]}"
        return True, ""
    # floating point: NaN and Inf placement identical, finite part within 64 eps in the 1-norm
    if not np.array_equal(np.isnan(ref), np.isnan(got)):
        return False, "NaN pattern differs"
    infr, infg = np.isinf(ref), np.isinf(got)
    if not np.array_equal(infr, infg) or not np.array_equal(ref[infr], got[infg]):
        return False, "Inf pattern differs"
    fin = np.isfinite(ref)
    r, g = ref[fin].astype(np.float64), got[fin].astype(np.float64)
    if add in ("MIN", "MAX"):
        # order-independent monoids: bit-exact apart from the sign of zero
        ok = np.array_equal(r, g)
        return ok, "" if ok else f"MIN/MAX values differ: max abs diff {np.abs(r - g).max()}"
    eps = np.finfo(grbref.NP_OF[type_]).eps
    err, nrm = np.abs(r - g).sum(), np.abs(r).sum()
    ok = err <= 64 * eps * nrm
    return ok, "" if ok else f"1-norm error {err} > 64 eps * {nrm}"


def compare(ref: dict, got: dict, add: str) -> tuple[bool, str]:
    for k in ("type", "nrows", "ncols", "nvals"):
        if ref[k] != got[k]:
            return False, f"{k}: ref {ref[k]} got {got[k]}"
    if not np.array_equal(ref["Ap"], got["Ap"]):
        return False, "vector pointers differ"
    if not np.array_equal(ref["Ai"], got["Ai"]):
        return False, "pattern differs"
    return values_equal(ref["type"], add, ref["Ax"], got["Ax"])


def run_mxm(G, gpu: bool, *, A, B, type_, semiring, M=None, mtype="BOOL", Cinit=None, accum=None,
            fmt="CSR", outp=GxB_DEFAULT, mask=GxB_DEFAULT, inp0=GxB_DEFAULT, inp1=GxB_DEFAULT,
            method=GxB_DEFAULT, ctype=None, cfmt=None, btype=None):
    """C<M> = accum (C, A*B) through GrB_mxm; returns the CSR export of C and the shim stats delta.
    btype: type of B if it differs from A's (the multiply then typecasts its operands)"""
    sr = semiring.replace("GxB_", "").split("_")
    ztype = "BOOL" if sr[1] in ("EQ", "NE", "GT", "LT", "GE", "LE") else sr[2]
    ctype = ctype or ztype
    a = import_sp(G, A, type_, fmt)
    b = a if B is A else import_sp(G, B, btype or type_, fmt)
    nrows = A.shape[1] if inp0 == GrB_TRAN else A.shape[0]
    ncols = B.shape[0] if inp1 == GrB_TRAN else B.shape[1]
    if Cinit is not None:
        c = import_sp(G, Cinit, ctype, cfmt or fmt)
    else:
        c = import_sp(G, sp.csr_matrix((nrows, ncols)), ctype, cfmt or fmt)
    m = import_sp(G, M, mtype, fmt) if M is not None else None
    d = G.descriptor(outp, mask, inp0, inp1, method)
    G.use_gpu(gpu)
    before = G.shim_stats()
    try:
        G.mxm(c, m, accum, semiring, a, b, d)
        G.matrix_nvals(c)
    finally:
        G.use_gpu(False)
    after = G.shim_stats()
    out = export_csr(G, c)
    G.matrix_free(a)
    if b is not a:
        G.matrix_free(b)
    if m is not None:
        G.matrix_free(m)
    G.descriptor_free(d)
    out["gpu_calls"] = after["gpu_calls"] - before["gpu_calls"]
    out["declined"] = after["declined"] - before["declined"]
    return out


def check_mxm(G, **kw):
    ref = run_mxm(G, False, **kw)
    got = run_mxm(G, not REF_ONLY, **kw)
    assert REF_ONLY or (got["gpu_calls"] >= 1 and got["declined"] == 0), "the GPU path did not run"
    add = kw["semiring"].replace("GxB_", "").split("_")[0]
    ok, why = compare(ref, got, add)
    assert ok, why
    return ref, got


def run_mv(G, gpu: bool, *, op, A, u, type_, semiring, n_out, mask=None, mtype="BOOL", winit=None,
           accum=None, fmt="CSR", outp=GxB_DEFAULT, maskd=GxB_DEFAULT, tran=GxB_DEFAULT,
           method=GxB_DEFAULT):
    """w<mask> = accum (w, A*u) (op='mxv') or accum (w, u'*A) (op='vxm').  u, mask, winit are
    (indices, values) pairs."""
    sr = semiring.replace("GxB_", "").split("_")
    ztype = "BOOL" if sr[1] in ("EQ", "NE", "GT", "LT", "GE", "LE") else sr[2]
    a = import_sp(G, A, type_, fmt)
    uv = G.vector_import(type_, u[0], u[1], u[2])
    if winit is not None:
        w = G.vector_import(ztype, n_out, winit[0], winit[1])
    else:
        w = G.vector_new(ztype, n_out)
    mv = G.vector_import(mtype, n_out, mask[0], mask[1]) if mask is not None else None
    if op == "mxv":
        d = G.descriptor(outp, maskd, tran, GxB_DEFAULT, method)
    else:
        d = G.descriptor(outp, maskd, GxB_DEFAULT, tran, method)
    G.use_gpu(gpu)
    before = G.shim_stats()
    try:
        if op == "mxv":
            G.mxv(w, mv, accum, semiring, a, uv, d)
        else:
            G.vxm(w, mv, accum, semiring, uv, a, d)
        G.vector_nvals(w)
    finally:
        G.use_gpu(False)
    after = G.shim_stats()
    out = G.vector_export(w)
    G.matrix_free(a)
    G.vector_free(uv)
    if mv is not None:
        G.vector_free(mv)
    G.descriptor_free(d)
    out["gpu_calls"] = after["gpu_calls"] - before["gpu_calls"]
    out["declined"] = after["declined"] - before["declined"]
    return out


def check_mv(G, **kw):
    ref = run_mv(G, False, **kw)
    got = run_mv(G, not REF_ONLY, **kw)
    assert REF_ONLY or (got["gpu_calls"] >= 1 and got["declined"] == 0), "the GPU path did not run"
    add = kw["semiring"].replace("GxB_", "").split("_")[0]
    assert ref["type"] == got["type"] and ref["n"] == got["n"]
    assert np.array_equal(ref["vi"], got["vi"]), "vector pattern differs"
    ok, why = values_equal(ref["type"], add, ref["vx"], got["vx"])
    assert ok, why
    return ref, got
