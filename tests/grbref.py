"""grbref.py -- TEST INFRASTRUCTURE: ctypes binding of the public GraphBLAS C API (GrB_* / GxB_*,
reference Include/GraphBLAS.h) of the compiled reference library oracle/_ref/libgraphblas_ref.so.

It plays two roles:
  * the oracle: with the shim disabled, GrB_mxm / GrB_mxv / GrB_vxm run the reference's own CPU
    path (GB_AxB_parallel -> Gustavson / heap / dot);
  * the unmodified caller: with the shim enabled, the *same* calls reach the B200 through the
    interposed GB_AxB_parallel (graphblas_b200/csrc/shim/gb_axb_parallel_shim.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` arm import it.
Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libgraphblas_ref.so")
DEMO_LIB = os.path.join(ROOT, "oracle", "_ref", "libgbdemo_ref.so")
SHIM_LIB = os.path.join(ROOT, "graphblas_b200", "libgb_b200_shim.so")
GPU_LIB = os.path.join(ROOT, "graphblas_b200", "libgb_b200.so")

# enums, reference Include/GraphBLAS.h:2784-2823
GrB_OUTP, GrB_MASK, GrB_INP0, GrB_INP1, GxB_AxB_METHOD = 0, 1, 2, 3, 1000
GxB_DEFAULT, GrB_REPLACE, GrB_SCMP, GrB_TRAN = 0, 1, 2, 3
GxB_AxB_GUSTAVSON, GxB_AxB_HEAP, GxB_AxB_DOT = 1001, 1002, 1003

NP_OF = {"BOOL": np.bool_, "INT8": np.int8, "UINT8": np.uint8, "INT16": np.int16,
         "UINT16": np.uint16, "INT32": np.int32, "UINT32": np.uint32, "INT64": np.int64,
         "UINT64": np.uint64, "FP32": np.float32, "FP64": np.float64}

_libc = C.CDLL(None)
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]
_libc.free.argtypes = [C.c_void_p]


class GrBError(RuntimeError):
    pass


class _MatrixHeader(C.Structure):
    """Leading fields of struct GB_Matrix_opaque (reference Source/Template/GB_matrix.h:193-208,
    307-315), read-only, to look at the raw T returned by GB_AxB_parallel."""
    _fields_ = [("magic", C.c_int64), ("type", C.c_void_p), ("type_size", C.c_size_t),
                ("hyper_ratio", C.c_double), ("plen", C.c_int64), ("vlen", C.c_int64),
                ("vdim", C.c_int64), ("nvec", C.c_int64), ("nvec_nonempty", C.c_int64),
                ("h", C.c_void_p), ("p", C.c_void_p), ("i", C.c_void_p), ("x", C.c_void_p),
                ("nzmax", C.c_int64), ("n_pending", C.c_int64), ("max_n_pending", C.c_int64),
                ("i_pending", C.c_void_p), ("j_pending", C.c_void_p), ("s_pending", C.c_void_p),
                ("type_pending", C.c_void_p), ("type_pending_size", C.c_size_t),
                ("operator_pending", C.c_void_p), ("nzombies", C.c_int64),
                ("AxB_method_used", C.c_int), ("queue_next", C.c_void_p),
                ("queue_prev", C.c_void_p), ("enqueued", C.c_bool), ("p_shallow", C.c_bool),
                ("h_shallow", C.c_bool), ("i_shallow", C.c_bool), ("x_shallow", C.c_bool),
                ("is_hyper", C.c_bool), ("is_csc", C.c_bool), ("sorted_pending", C.c_bool)]


def available() -> bool:
    return os.path.exists(REF_LIB)


class GraphBLAS:
    """One process-wide instance: loads the shim (optional) ahead of the reference library."""
    _inst = None

    @classmethod
    def get(cls, with_shim: bool, pinned: bool = False) -> "GraphBLAS":
        if cls._inst is None:
            cls._inst = cls(with_shim, pinned)
        if with_shim and cls._inst.shim is None:
            raise GrBError("GraphBLAS was already loaded without the shim in this process")
        return cls._inst

    def __init__(self, with_shim: bool, pinned: bool = False):
        """pinned: start the reference with GxB_init (mode, gb200_host_malloc, gb200_host_calloc,
        gb200_host_realloc, gb200_host_free) -- reference Include/GraphBLAS.h:330-340 -- which is what
        INTEGRATION.md tells a host application to do: every GraphBLAS array (operands, and the T the
        shim builds with GB_create) then lives in page-locked memory."""
        if not available():
            raise GrBError(f"{REF_LIB} missing: run `make -C oracle -f Makefile.ref` where "
                           "/root/reference exists")
        self.shim = None
        if with_shim:
            if not os.path.exists(SHIM_LIB):
                raise GrBError(f"{SHIM_LIB} missing: run `make -C graphblas_b200 shim`")
            # order matters: the shim's GB_AxB_parallel must come first in the global scope
            self.shim = C.CDLL(SHIM_LIB, mode=C.RTLD_GLOBAL)
            self.shim.gb200_shim_enable(0)
        self.lib = C.CDLL(REF_LIB, mode=C.RTLD_GLOBAL)
        self.lib.GrB_error.restype = C.c_char_p
        self.host_malloc = None
        self._free = _libc.free
        if pinned:
            gpu = C.CDLL(GPU_LIB, mode=C.RTLD_GLOBAL)
            gpu.gb200_host_malloc.restype = C.c_void_p
            gpu.gb200_host_malloc.argtypes = [C.c_size_t]
            self.host_malloc = gpu.gb200_host_malloc
            gpu.gb200_host_free.restype = None
            gpu.gb200_host_free.argtypes = [C.c_void_p]
            self._free = gpu.gb200_host_free
            fns = [C.cast(getattr(gpu, "gb200_host_" + f), C.c_void_p)
                   for f in ("malloc", "calloc", "realloc", "free")]
            self.ok(self.lib.GxB_init(0, *fns), "GxB_init")
        else:
            self.ok(self.lib.GrB_init(0), "GrB_init")        # GrB_NONBLOCKING

    # ---- plumbing ---------------------------------------------------------------------------
    def ok(self, info: int, where: str) -> None:
        if info != 0:
            raise GrBError(f"{where} -> GrB_Info {info}: {self.lib.GrB_error().decode()}")

    def obj(self, name: str) -> C.c_void_p:
        """A predefined object: GrB_FP64, GxB_PLUS_TIMES_FP64, GrB_PLUS_FP64, GxB_PLUS_INT64_MONOID."""
        return C.c_void_p.in_dll(self.lib, name)

    def use_gpu(self, on: bool) -> None:
        if self.shim is None:
            if on:
                raise GrBError("shim not loaded")
            return
        self.shim.gb200_shim_enable(1 if on else 0)

    def shim_stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.shim.gb200_shim_stats(C.byref(a), C.byref(b), C.byref(c))
        return {"gpu_calls": a.value, "forwarded": b.value, "declined": c.value}

    def shim_cache(self, on=None):
        """operand residency of the library behind the shim: on = True / False switches it, None only
        reads the counters -> dict(hits, misses, invalidations)"""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.shim.gb200_shim_cache(C.c_int(-1 if on is None else int(bool(on))), C.byref(a), C.byref(b),
                                   C.byref(c))
        return {"hits": a.value, "misses": b.value, "invalidations": c.value}

    def shim_last(self):
        ms, fl = C.c_double(), C.c_int64()
        self.shim.gb200_shim_last(C.byref(ms), C.byref(fl))
        return ms.value, fl.value

    def _malloc_copy(self, a: np.ndarray) -> C.c_void_p:
        """a copy of `a` in memory from the allocator GraphBLAS was started with (import hands
        ownership of the arrays to the library, which later frees them with that allocator)"""
        n = max(a.nbytes, 8)
        p = self.host_malloc(n) if self.host_malloc is not None else _libc.malloc(n)
        if a.nbytes:
            C.memmove(p, a.ctypes.data, a.nbytes)
        return C.c_void_p(p)

    # ---- import / export (reference Include/GraphBLAS.h:5751-6064) -----------------------------
    def matrix_import(self, fmt: str, type_: str, nrows: int, ncols: int, Ap, Ai, Ax, Ah=None):
        """fmt in CSR, CSC, HyperCSR, HyperCSC.  Arrays are copied into malloc'd memory whose
        ownership moves to GraphBLAS."""
        Ap = np.ascontiguousarray(Ap, dtype=np.uint64)
        Ai = np.ascontiguousarray(Ai, dtype=np.uint64)
        Ax = np.ascontiguousarray(Ax, dtype=NP_OF[type_])
        nvals = int(Ap[-1])
        A = C.c_void_p()
        pp, pi, px = self._malloc_copy(Ap), self._malloc_copy(Ai), self._malloc_copy(Ax)
        t = self.obj("GrB_" + type_)
        if fmt in ("CSR", "CSC"):
            fn = getattr(self.lib, "GxB_Matrix_import_" + fmt)
            info = fn(C.byref(A), t, C.c_uint64(nrows), C.c_uint64(ncols), C.c_uint64(nvals),
                      C.c_int64(-1), C.byref(pp), C.byref(pi), C.byref(px), None)
        else:
            Ah = np.ascontiguousarray(Ah, dtype=np.uint64)
            ph = self._malloc_copy(Ah)
            fn = getattr(self.lib, "GxB_Matrix_import_" + fmt)
            info = fn(C.byref(A), t, C.c_uint64(nrows), C.c_uint64(ncols), C.c_uint64(nvals),
                      C.c_int64(-1), C.c_uint64(len(Ah)), C.byref(ph), C.byref(pp), C.byref(pi),
                      C.byref(px), None)
        self.ok(info, "GxB_Matrix_import_" + fmt)
        return A

    def matrix_export(self, A: C.c_void_p, fmt: str):
        """Destructive export; returns dict(type, nrows, ncols, nvals, Ap, Ai, Ax[, Ah])."""
        t = C.c_void_p()
        nrows, ncols, nvals = C.c_uint64(), C.c_uint64(), C.c_uint64()
        nonempty = C.c_int64()
        pp, pi, px, ph = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        nvec = C.c_uint64()
        fn = getattr(self.lib, "GxB_Matrix_export_" + fmt)
        if fmt in ("CSR", "CSC"):
            info = fn(C.byref(A), C.byref(t), C.byref(nrows), C.byref(ncols), C.byref(nvals),
                      C.byref(nonempty), C.byref(pp), C.byref(pi), C.byref(px), None)
            np_len = (nrows.value if fmt == "CSR" else ncols.value) + 1
        else:
            info = fn(C.byref(A), C.byref(t), C.byref(nrows), C.byref(ncols), C.byref(nvals),
                      C.byref(nonempty), C.byref(nvec), C.byref(ph), C.byref(pp), C.byref(pi),
                      C.byref(px), None)
            np_len = nvec.value + 1
        self.ok(info, "GxB_Matrix_export_" + fmt)
        tname = self.type_name(t)
        out = {"type": tname, "nrows": nrows.value, "ncols": ncols.value, "nvals": nvals.value,
               "nonempty": nonempty.value}
        out["Ap"] = self._take(pp, np_len, np.int64)
        out["Ai"] = self._take(pi, nvals.value, np.int64)
        out["Ax"] = self._take(px, nvals.value, NP_OF[tname])
        if fmt not in ("CSR", "CSC"):
            out["Ah"] = self._take(ph, nvec.value, np.int64)
        return out

    def _take(self, ptr: C.c_void_p, n: int, dt) -> np.ndarray:
        dt = np.dtype(dt)
        if not ptr.value or n == 0:
            if ptr.value:
                self._free(ptr)
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (n * dt.itemsize)).from_address(ptr.value)
        a = np.frombuffer(buf, dtype=dt, count=n).copy()
        self._free(ptr)
        return a

    def type_name(self, t: C.c_void_p) -> str:
        for name in NP_OF:
            if self.obj("GrB_" + name).value == t.value:
                return name
        raise GrBError("unknown type handle")

    def vector_import(self, type_: str, n: int, vi, vx):
        vi = np.ascontiguousarray(vi, dtype=np.uint64)
        vx = np.ascontiguousarray(vx, dtype=NP_OF[type_])
        v = C.c_void_p()
        pi, px = self._malloc_copy(vi), self._malloc_copy(vx)
        self.ok(self.lib.GxB_Vector_import(C.byref(v), self.obj("GrB_" + type_), C.c_uint64(n),
                                           C.c_uint64(len(vi)), C.byref(pi), C.byref(px), None),
                "GxB_Vector_import")
        return v

    def vector_export(self, v: C.c_void_p):
        t = C.c_void_p()
        n, nvals = C.c_uint64(), C.c_uint64()
        pi, px = C.c_void_p(), C.c_void_p()
        self.ok(self.lib.GxB_Vector_export(C.byref(v), C.byref(t), C.byref(n), C.byref(nvals),
                                           C.byref(pi), C.byref(px), None), "GxB_Vector_export")
        tname = self.type_name(t)
        return {"type": tname, "n": n.value, "nvals": nvals.value,
                "vi": self._take(pi, nvals.value, np.int64),
                "vx": self._take(px, nvals.value, NP_OF[tname])}

    # ---- objects ------------------------------------------------------------------------------
    def matrix_new(self, type_: str, nrows: int, ncols: int):
        A = C.c_void_p()
        self.ok(self.lib.GrB_Matrix_new(C.byref(A), self.obj("GrB_" + type_), C.c_uint64(nrows),
                                        C.c_uint64(ncols)), "GrB_Matrix_new")
        return A

    def vector_new(self, type_: str, n: int):
        v = C.c_void_p()
        self.ok(self.lib.GrB_Vector_new(C.byref(v), self.obj("GrB_" + type_), C.c_uint64(n)),
                "GrB_Vector_new")
        return v

    def matrix_dup(self, A):
        B = C.c_void_p()
        self.ok(self.lib.GrB_Matrix_dup(C.byref(B), A), "GrB_Matrix_dup")
        return B

    def vector_dup(self, v):
        w = C.c_void_p()
        self.ok(self.lib.GrB_Vector_dup(C.byref(w), v), "GrB_Vector_dup")
        return w

    def matrix_free(self, A):
        self.lib.GrB_Matrix_free(C.byref(A))

    def vector_free(self, v):
        self.lib.GrB_Vector_free(C.byref(v))

    def demo_random_matrix(self, nrows: int, ncols: int, ntuples: int, seed=None):
        """the reference's own generator (Demo/Source/random_matrix.c:26-181 over the LCG of
        Demo/Source/simple_rand.c:31-64), from the compiled Demo library: a GrB_FP64 matrix.
        seed: simple_rand_seed (seed) first (None: continue the stream, as consecutive calls in one
        program do)."""
        if not hasattr(self, "demo"):
            self.demo = C.CDLL(DEMO_LIB, mode=C.RTLD_GLOBAL)
        if seed is not None:
            self.demo.simple_rand_seed(C.c_uint64(seed))
        A = C.c_void_p()
        self.ok(self.demo.random_matrix(C.byref(A), C.c_bool(False), C.c_bool(False), C.c_int64(nrows),
                                        C.c_int64(ncols), C.c_int64(ntuples), C.c_int(1), C.c_bool(False)),
                "random_matrix")
        return A

    def matrix_set_element(self, A, type_: str, i: int, j: int, x) -> None:
        """GrB_Matrix_setElement_<type> (Include/GraphBLAS.h): A(i,j) = x"""
        ct = {"FP64": C.c_double, "FP32": C.c_float, "INT64": C.c_int64, "INT32": C.c_int32,
              "UINT64": C.c_uint64, "UINT32": C.c_uint32, "BOOL": C.c_bool}[type_]
        fn = getattr(self.lib, "GrB_Matrix_setElement_" + type_)
        fn.argtypes = [C.c_void_p, ct, C.c_uint64, C.c_uint64]
        self.ok(fn(A, ct(x), C.c_uint64(i), C.c_uint64(j)), "GrB_Matrix_setElement")

    def matrix_nvals(self, A) -> int:
        n = C.c_uint64()
        self.ok(self.lib.GrB_Matrix_nvals(C.byref(n), A), "GrB_Matrix_nvals")
        return n.value

    def vector_nvals(self, v) -> int:
        n = C.c_uint64()
        self.ok(self.lib.GrB_Vector_nvals(C.byref(n), v), "GrB_Vector_nvals")
        return n.value

    def descriptor(self, outp=GxB_DEFAULT, mask=GxB_DEFAULT, inp0=GxB_DEFAULT, inp1=GxB_DEFAULT,
                   method=GxB_DEFAULT):
        if (outp, mask, inp0, inp1, method) == (0, 0, 0, 0, 0):
            return None
        d = C.c_void_p()
        self.ok(self.lib.GrB_Descriptor_new(C.byref(d)), "GrB_Descriptor_new")
        for f, v in ((GrB_OUTP, outp), (GrB_MASK, mask), (GrB_INP0, inp0), (GrB_INP1, inp1),
                     (GxB_AxB_METHOD, method)):
            if v != GxB_DEFAULT:
                self.ok(self.lib.GrB_Descriptor_set(d, f, v), "GrB_Descriptor_set")
        return d

    def descriptor_free(self, d):
        if d is not None:
            self.lib.GrB_Descriptor_free(C.byref(d))

    # ---- the path under test --------------------------------------------------------------------
    def mxm(self, Cm, M, accum, semiring: str, A, B, desc):
        self.ok(self.lib.GrB_mxm(Cm, M, self.obj(accum) if accum else None, self.obj(semiring), A, B,
                                 desc), "GrB_mxm")

    def mxv(self, w, mask, accum, semiring: str, A, u, desc):
        self.ok(self.lib.GrB_mxv(w, mask, self.obj(accum) if accum else None, self.obj(semiring), A,
                                 u, desc), "GrB_mxv")

    def vxm(self, w, mask, accum, semiring: str, u, A, desc):
        self.ok(self.lib.GrB_vxm(w, mask, self.obj(accum) if accum else None, self.obj(semiring), u,
                                 A, desc), "GrB_vxm")

    def select(self, Cm, M, accum, op: str, A, k: int, desc):
        """GxB_select (Include/GraphBLAS.h:4670-4679) with a built-in operator: op in TRIL, TRIU, DIAG,
        OFFDIAG, NONZERO"""
        kk = C.c_int64(k)
        self.ok(self.lib.GxB_Matrix_select(Cm, M, self.obj(accum) if accum else None,
                                           self.obj("GxB_" + op), A, C.byref(kk), desc),
                "GxB_select")

    def matrix_reduce(self, A, type_: str, monoid: str, accum=None, init=0):
        """GrB_Matrix_reduce_<type> (&c, accum, monoid, A, NULL) with c = init: the scalar c"""
        ct = {"FP64": C.c_double, "FP32": C.c_float, "INT64": C.c_int64, "INT32": C.c_int32,
              "UINT64": C.c_uint64, "UINT32": C.c_uint32, "INT16": C.c_int16, "UINT16": C.c_uint16,
              "INT8": C.c_int8, "UINT8": C.c_uint8, "BOOL": C.c_bool}[type_]
        c = ct(init)
        fn = getattr(self.lib, "GrB_Matrix_reduce_" + type_)
        self.ok(fn(C.byref(c), self.obj(accum) if accum else None, self.obj(monoid), A, None),
                "GrB_Matrix_reduce")
        return c.value

    def shim_reduce_calls(self) -> int:
        self.shim.gb200_shim_reduce_calls.restype = C.c_int64
        return self.shim.gb200_shim_reduce_calls()

    def shim_select_calls(self) -> int:
        self.shim.gb200_shim_select_calls.restype = C.c_int64
        return self.shim.gb200_shim_select_calls()

    def raw(self, T) -> dict:
        """the raw content of a GrB_Matrix in the CSC-agnostic layout (header fields, p, h, i, x)"""
        hdr = _MatrixHeader.from_address(T.value)
        out = {"vlen": hdr.vlen, "vdim": hdr.vdim, "nvec": hdr.nvec, "is_hyper": bool(hdr.is_hyper),
               "nvec_nonempty": hdr.nvec_nonempty}
        nnz = 0
        if hdr.nzmax > 0:
            p = np.ctypeslib.as_array((C.c_int64 * (hdr.nvec + 1)).from_address(hdr.p)).copy()
            nnz = int(p[-1])
        else:
            p = np.zeros(hdr.nvec + 1, dtype=np.int64)
        out["p"] = p
        out["h"] = (np.ctypeslib.as_array((C.c_int64 * hdr.nvec).from_address(hdr.h)).copy()
                    if hdr.is_hyper and hdr.nvec > 0 else
                    (np.zeros(0, dtype=np.int64) if hdr.is_hyper else None))
        tname = self.type_name(C.c_void_p(hdr.type))
        dt = np.dtype(NP_OF[tname])
        out["type"] = tname
        out["i"] = (np.ctypeslib.as_array((C.c_int64 * nnz).from_address(hdr.i)).copy()
                    if nnz else np.zeros(0, dtype=np.int64))
        out["x"] = (np.frombuffer((C.c_char * (nnz * dt.itemsize)).from_address(hdr.x), dtype=dt,
                                  count=nnz).copy() if nnz else np.zeros(0, dtype=dt))
        return out

    # ---- GB_transpose (Source/GB.h:2153-2162), found through the GLOBAL symbol scope -------------
    def seam_transpose(self, A, ctype, c_is_csc: bool):
        """GB_transpose (&T, ctype, C_is_csc, A, NULL, NULL) as the reference's own callers reach it: the
        shim's interposer when the shim is loaded (it forwards to the reference while switched off), the
        reference's otherwise.  Returns the raw T."""
        T = C.c_void_p()
        fn = C.CDLL(None).GB_transpose
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_bool, C.c_void_p, C.c_void_p, C.c_void_p]
        self.ok(fn(C.byref(T), self.obj("GrB_" + ctype) if ctype else None, c_is_csc, A, None, None),
                "GB_transpose")
        out = self.raw(T)
        out["is_csc"] = bool(_MatrixHeader.from_address(T.value).is_csc)
        self.matrix_free(T)
        return out

    def seam_accum_mask(self, Cm, M, accum, T, replace: bool, comp: bool) -> dict:
        """GB_accum_mask (C, M, NULL, accum, &T, C_replace, Mask_comp, NULL) (Source/GB_accum_mask.c:130) found
        through the global symbol scope (the shim's interposer when loaded; it forwards while switched off).
        T is consumed.  Returns the raw C after its pending work was assembled."""
        fn = C.CDLL(None).GB_accum_mask
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_bool, C.c_bool, C.c_void_p]
        th = C.c_void_p(T.value)
        self.ok(fn(Cm, M, None, self.obj(accum) if accum else None, C.byref(th), replace, comp, None),
                "GB_accum_mask")
        self.matrix_nvals(Cm)
        return self.raw(Cm)

    def assign_scalar(self, Cm, M, accum, type_: str, x, desc, nrows: int, ncols=None):
        """GrB_Matrix_assign_<type> (C, Mask, accum, x, GrB_ALL, nrows, GrB_ALL, ncols, desc), or
        GrB_Vector_assign_<type> (w, mask, accum, x, GrB_ALL, n, desc) when ncols is None
        (Include/GraphBLAS.h:4260, 4477; Demo/Source/bfs5m.c:74)"""
        ct = {"FP64": C.c_double, "FP32": C.c_float, "INT64": C.c_int64, "INT32": C.c_int32,
              "UINT64": C.c_uint64, "UINT32": C.c_uint32, "INT16": C.c_int16, "UINT16": C.c_uint16,
              "INT8": C.c_int8, "UINT8": C.c_uint8, "BOOL": C.c_bool}[type_]
        all_ = C.c_void_p.in_dll(self.lib, "GrB_ALL")
        acc = self.obj(accum) if accum else None
        if ncols is None:
            fn = getattr(self.lib, "GrB_Vector_assign_" + type_)
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, ct, C.c_void_p, C.c_uint64, C.c_void_p]
            self.ok(fn(Cm, M, acc, x, all_, nrows, desc), "GrB_Vector_assign")
        else:
            fn = getattr(self.lib, "GrB_Matrix_assign_" + type_)
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, ct, C.c_void_p, C.c_uint64, C.c_void_p,
                           C.c_uint64, C.c_void_p]
            self.ok(fn(Cm, M, acc, x, all_, nrows, all_, ncols, desc), "GrB_Matrix_assign")

    def shim_assign_calls(self) -> int:
        self.shim.gb200_shim_assign_calls.restype = C.c_int64
        return self.shim.gb200_shim_assign_calls()

    def transpose(self, Cm, M, accum, A, desc):
        """GrB_transpose (Include/GraphBLAS.h): C<M> = accum (C, A')"""
        self.ok(self.lib.GrB_transpose(Cm, M, self.obj(accum) if accum else None, A, desc), "GrB_transpose")

    def shim_transpose_calls(self) -> int:
        self.shim.gb200_shim_transpose_calls.restype = C.c_int64
        return self.shim.gb200_shim_transpose_calls()

    def shim_accum_mask_calls(self) -> int:
        self.shim.gb200_shim_accum_mask_calls.restype = C.c_int64
        return self.shim.gb200_shim_accum_mask_calls()

    def shim_accum_mask_min(self, nnz: int) -> None:
        """fewer entries in C and T together: the host's own GB_accum_mask; < 0: the default rule (65536 while
        the residency cache is on, never while it is off)"""
        self.shim.gb200_shim_accum_mask_min.argtypes = [C.c_int64]
        self.shim.gb200_shim_accum_mask_min(nnz)

    def shim_transpose_min(self, nnz: int) -> None:
        """matrices with fewer entries are transposed by the host (default 65536)"""
        self.shim.gb200_shim_transpose_min.argtypes = [C.c_int64]
        self.shim.gb200_shim_transpose_min(nnz)

    # ---- the seam itself: the reference's own GB_AxB_parallel (Source/GB.h:1522-1537) ---------
    def seam_axb(self, M, mask_comp: bool, A, B, semiring: str, flipxy: bool, do_adotb: bool,
                 method: int = GxB_DEFAULT):
        """Calls the reference library's internal GB_AxB_parallel directly on GrB_Matrix handles and
        returns (T as dict in the CSC-agnostic layout, method_used, mask_applied).  M, A, B must be
        stored by column (import_CSC / import_HyperCSC) so that the handle's vectors are exactly the
        seam's vectors."""
        T = C.c_void_p()
        used = C.c_int(0)
        applied = C.c_bool(False)
        fn = self.lib.GB_AxB_parallel
        fn.restype = C.c_int
        info = fn(C.byref(T), M, C.c_bool(mask_comp), A, B, self.obj(semiring), C.c_bool(flipxy),
                  C.c_bool(do_adotb), C.c_int(method), C.byref(used), C.byref(applied), None)
        self.ok(info, "GB_AxB_parallel")
        out = self.raw(T)
        self.matrix_free(T)
        return out, used.value, bool(applied.value)

    def reduce_int64(self, A, monoid="GxB_PLUS_INT64_MONOID") -> int:
        s = C.c_int64(0)
        self.ok(self.lib.GrB_Matrix_reduce_INT64(C.byref(s), None, self.obj(monoid), A, None),
                "GrB_Matrix_reduce_INT64")
        return s.value
