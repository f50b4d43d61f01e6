"""The reference-side binding on a box without a GPU: the shim loads ahead of the reference library,
stays out of the way while disabled, and DECLINES what the device cannot compute (user-defined
operators are host function pointers) -- loudly by default, by delegation when asked to."""
import numpy as np
import scipy.sparse as sp

import gen


# ---------------------------------------------------------------------------------------------
# decline rule: a user-defined operator is a host function pointer and cannot run on
# the device.  Default: the seam fails loudly (GrB_PANIC); GB200_SHIM_FORWARD=1 delegates the call to
# the host library, and the result is the reference's.
# ---------------------------------------------------------------------------------------------
def test_user_defined_operator_is_declined(G, monkeypatch):
    import ctypes as C
    FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))

    def mult(z, x, y):
        z[0] = x[0] * y[0] + 1.0
    cb = FN(mult)
    op, mon, sr = C.c_void_p(), C.c_void_p(), C.c_void_p()
    fp64 = G.obj("GrB_FP64")
    G.ok(G.lib.GB_BinaryOp_new(C.byref(op), cb, fp64, fp64, fp64, b"mult_plus_one"), "GB_BinaryOp_new")
    G.ok(G.lib.GrB_Monoid_new_FP64(C.byref(mon), G.obj("GrB_PLUS_FP64"), C.c_double(0.0)), "GrB_Monoid_new")
    G.ok(G.lib.GrB_Semiring_new(C.byref(sr), mon, op), "GrB_Semiring_new")
    A = gen.er(60, 50, 400, 41)
    B = gen.er(50, 40, 350, 42)

    def run(gpu):
        from parity import import_sp, export_csr
        a, b = import_sp(G, A, "FP64", "CSR"), import_sp(G, B, "FP64", "CSR")
        c = import_sp(G, sp.csr_matrix((60, 40)), "FP64", "CSR")
        G.use_gpu(gpu)
        try:
            info = G.lib.GrB_mxm(c, None, None, sr, a, b, None)
            if info == 0:
                G.matrix_nvals(c)
        finally:
            G.use_gpu(False)
        out = export_csr(G, c) if info == 0 else None
        G.matrix_free(a)
        G.matrix_free(b)
        return info, out

    info_ref, ref = run(False)
    assert info_ref == 0
    before = G.shim_stats()
    monkeypatch.delenv("GB200_SHIM_FORWARD", raising=False)
    info, _ = run(True)
    after = G.shim_stats()
    assert info != 0, "a user-defined operator must not be computed silently"
    assert after["declined"] == before["declined"] + 1 and after["gpu_calls"] == before["gpu_calls"]
    monkeypatch.setenv("GB200_SHIM_FORWARD", "1")
    info, got = run(True)
    final = G.shim_stats()
    assert info == 0 and final["forwarded"] == after["forwarded"] + 1
    assert np.array_equal(ref["Ap"], got["Ap"]) and np.array_equal(ref["Ai"], got["Ai"])
    assert np.array_equal(ref["Ax"], got["Ax"])


# ---------------------------------------------------------------------------------------------
# The neighbours of the multiply (GB_transpose, GB_accum_mask, GB_select, GB_reduce_to_scalar) follow the
# same rule: a call the shim took and the device FAILED on (here: no device at all) fails loudly; it is
# never computed on the host behind the caller's back unless GB200_SHIM_FORWARD=1 asks for that.
# ---------------------------------------------------------------------------------------------
def test_neighbours_fail_loudly_without_a_device(G, has_gpu, monkeypatch, capfd):
    import ctypes as C
    import pytest
    from parity import import_sp, export_csr, compare
    if has_gpu:
        pytest.skip("needs a box without a GPU")
    A = gen.er(90, 70, 900, 51)

    def stats():
        f, w = C.c_int64(), C.c_int64()
        G.shim.gb200_shim_neighbour_stats(C.byref(f), C.byref(w))
        return f.value, w.value

    def run(gpu):
        a = import_sp(G, A, "FP64", "CSR")
        c = import_sp(G, sp.csr_matrix((70, 90)), "FP64", "CSR")
        G.shim_transpose_min(0)
        G.use_gpu(gpu)
        try:
            info = G.lib.GrB_transpose(c, None, None, a, None)
            if info == 0:
                G.matrix_nvals(c)
        finally:
            G.use_gpu(False)
            G.shim_transpose_min(65536)
        out = export_csr(G, c) if info == 0 else None
        G.matrix_free(a)
        return info, out

    info_ref, ref = run(False)
    assert info_ref == 0
    monkeypatch.delenv("GB200_SHIM_FORWARD", raising=False)
    f0, w0 = stats()
    info, _ = run(True)
    f1, w1 = stats()
    assert info != 0, "a transpose the device failed on must not be computed silently on the host"
    assert f1 == f0 + 1 and w1 == w0
    assert "GB_transpose failed on the device" in capfd.readouterr().err
    monkeypatch.setenv("GB200_SHIM_FORWARD", "1")
    info, got = run(True)
    f2, w2 = stats()
    assert info == 0 and w2 == w1 + 1
    ok, why = compare(ref, got, "MIN")
    assert ok, why
