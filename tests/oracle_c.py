"""oracle_c.py -- TEST INFRASTRUCTURE: ctypes binding of oracle/libgb_oracle.so, the plain-C CPU
restatement of the reference's multiply at the GB_AxB_parallel seam (oracle/gb_oracle.c).  Used as a
checker only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

import graphblas_b200 as gb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "libgb_oracle.so")
_lib = None


class _OMat(C.Structure):
    _fields_ = [("vlen", C.c_int64), ("vdim", C.c_int64), ("nvec", C.c_int64), ("p", C.c_void_p),
                ("h", C.c_void_p), ("i", C.c_void_p), ("x", C.c_void_p), ("type_code", C.c_int32),
                ("is_hyper", C.c_int32)]


class _ORes(C.Structure):
    _fields_ = [("vlen", C.c_int64), ("vdim", C.c_int64), ("nvec", C.c_int64),
                ("nvec_nonempty", C.c_int64), ("nnz", C.c_int64), ("p", C.c_void_p),
                ("h", C.c_void_p), ("i", C.c_void_p), ("x", C.c_void_p), ("is_hyper", C.c_int32),
                ("type_code", C.c_int32), ("mask_applied", C.c_int32), ("method_used", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} missing: make -C oracle")
        _lib = C.CDLL(LIB)
        _lib.oracle_flopcount.restype = C.c_int64
    return _lib


def _om(m: gb.Matrix) -> _OMat:
    return _OMat(m.vlen, m.vdim, m.nvec, m.p.ctypes.data,
                 m.h.ctypes.data if m.h is not None else None,
                 m.i.ctypes.data if m.i.size else None, m.x.ctypes.data if m.x.size else None,
                 gb.TYPES[m.type][0], 1 if m.h is not None else 0)


def _copy(ptr, n, dt):
    dt = np.dtype(dt)
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dt)
    return np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(ptr), dtype=dt, count=n).copy()


def axb(M, mask_comp, A, B, semiring: gb.Semiring, do_adotb=False, info=None) -> gb.Matrix:
    r = _ORes()
    cm = _om(M) if M is not None else None
    ca, cb = _om(A), _om(B)
    s = semiring.c()
    rc = lib().oracle_AxB(C.byref(r), C.byref(cm) if cm is not None else None, int(mask_comp),
                          C.byref(ca), C.byref(cb), s.add_opcode, s.mult_opcode, s.xy_code, s.z_code,
                          s.flipxy, int(do_adotb))
    if rc != 0:
        raise RuntimeError(f"oracle_AxB failed: {rc}")
    tname, dt = gb.TYPE_BY_CODE[r.type_code]
    out = gb.Matrix(r.vlen, r.vdim, _copy(r.p, r.nvec + 1, np.int64), _copy(r.i, r.nnz, np.int64),
                    _copy(r.x, r.nnz, dt), _copy(r.h, r.nvec, np.int64) if r.is_hyper else None, tname)
    if info is not None:
        info.update(mask_applied=r.mask_applied, method_used=r.method_used,
                    nvec_nonempty=r.nvec_nonempty, is_hyper=r.is_hyper)
    lib().oracle_free(C.byref(r))
    return out


def transpose(A, ctype=None, hyper_ratio=0.0625, info=None) -> gb.Matrix:
    """GB_transpose (&T, ctype, csc, A, NULL) restated: T in the form the reference ends with"""
    r = _ORes()
    ca = _om(A)
    code = gb.TYPES[ctype][0] if ctype is not None else ca.type_code
    lib().oracle_transpose.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    rc = lib().oracle_transpose(C.byref(r), C.byref(ca), code, hyper_ratio)
    if rc != 0:
        raise RuntimeError(f"oracle_transpose failed: {rc}")
    tname, dt = gb.TYPE_BY_CODE[r.type_code]
    out = gb.Matrix(r.vlen, r.vdim, _copy(r.p, r.nvec + 1, np.int64), _copy(r.i, r.nnz, np.int64),
                    _copy(r.x, r.nnz, dt), _copy(r.h, r.nvec, np.int64) if r.is_hyper else None, tname)
    if info is not None:
        info.update(nvec_nonempty=r.nvec_nonempty, is_hyper=r.is_hyper)
    lib().oracle_free(C.byref(r))
    return out


def accum_mask(Cm, T, M=None, mask_comp=False, replace=False, accum=None, hyper=False) -> gb.Matrix:
    """GB_accum_mask (C, M, NULL, accum, &T, C_replace, Mask_comp) restated: the new C.
    accum = (operator name, type name of its inputs) or None"""
    r = _ORes()
    cc, ct = _om(Cm), _om(T)
    cm = _om(M) if M is not None else None
    op, xy = (gb.OPCODES[accum[0]], gb.TYPES[accum[1]][0]) if accum is not None else (0, 0)
    rc = lib().oracle_accum_mask(C.byref(r), C.byref(cc), C.byref(ct), C.byref(cm) if cm is not None else None,
                                 int(mask_comp), int(replace), op, xy, int(hyper))
    if rc != 0:
        raise RuntimeError(f"oracle_accum_mask failed: {rc}")
    tname, dt = gb.TYPE_BY_CODE[r.type_code]
    out = gb.Matrix(r.vlen, r.vdim, _copy(r.p, r.nvec + 1, np.int64), _copy(r.i, r.nnz, np.int64),
                    _copy(r.x, r.nnz, dt), _copy(r.h, r.nvec, np.int64) if r.is_hyper else None, tname)
    lib().oracle_free(C.byref(r))
    return out


def flopcount(M, A, B):
    out = np.empty(B.nvec + 1, dtype=np.int64)
    cm = _om(M) if M is not None else None
    ca, cb = _om(A), _om(B)
    total = lib().oracle_flopcount(C.byref(cm) if cm is not None else None, C.byref(ca), C.byref(cb),
                                   out.ctypes.data_as(C.c_void_p))
    return out, total
