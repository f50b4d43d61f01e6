"""graphblas_b200 -- thin ctypes binding of libgb_b200.so (include/gb_b200.h).

The product is the CUDA library and its C ABI; this module only exists so that the tests and
bench.py can drive that ABI from Python with numpy arrays.  It never computes anything itself and it
raises at import time if the CUDA library has not been built (`make -C graphblas_b200`, or
`python -c "import __graft_entry__ as g; g.build()"`).

Vocabulary follows the reference (SuiteSparse:GraphBLAS v2.3.3): a matrix is `vdim` sparse vectors
of length `vlen` ("CSC-agnostic", Source/Template/GB_matrix.h:193-208); `h` is the hyperlist.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgb_b200.so")
SHIM_PATH = os.path.join(_HERE, "libgb_b200_shim.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA library first (make -C graphblas_b200). "
        "There is no CPU fallback.")

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

# ---- codes (identical to the reference's GB_Type_code / GB_Opcode, Source/GB.h:450-550) ----------
TYPES = {
    "BOOL": (0, np.bool_), "INT8": (1, np.int8), "UINT8": (2, np.uint8), "INT16": (3, np.int16),
    "UINT16": (4, np.uint16), "INT32": (5, np.int32), "UINT32": (6, np.uint32),
    "INT64": (7, np.int64), "UINT64": (8, np.uint64), "FP32": (9, np.float32),
    "FP64": (10, np.float64),
}
TYPE_BY_CODE = {v[0]: (k, v[1]) for k, v in TYPES.items()}
OPCODES = {
    "FIRST": 7, "SECOND": 8, "MIN": 9, "MAX": 10, "PLUS": 11, "MINUS": 12, "TIMES": 13, "DIV": 14,
    "ISEQ": 15, "ISNE": 16, "ISGT": 17, "ISLT": 18, "ISGE": 19, "ISLE": 20,
    "LOR": 21, "LAND": 22, "LXOR": 23, "EQ": 24, "NE": 25, "GT": 26, "LT": 27, "GE": 28, "LE": 29,
}
COMPARE_OPS = ("EQ", "NE", "GT", "LT", "GE", "LE")
METHOD_DEFAULT, METHOD_GUSTAVSON, METHOD_HEAP, METHOD_DOT = 0, 1001, 1002, 1003

STATUS = {0: "SUCCESS", 1: "OUT_OF_MEMORY", 2: "NOT_SUPPORTED", 3: "INVALID", 4: "NO_DEVICE",
          5: "CUDA_ERROR"}


class GB200Error(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib.gb200_last_error().decode(errors="replace")
        super().__init__(f"{where}: GB200_{STATUS.get(code, code)}: {msg}")


class _CMatrix(C.Structure):
    _fields_ = [("vlen", C.c_int64), ("vdim", C.c_int64), ("nvec", C.c_int64),
                ("p", C.c_void_p), ("h", C.c_void_p), ("i", C.c_void_p), ("x", C.c_void_p),
                ("type_code", C.c_int32), ("reserved", C.c_int32)]


class _CSemiring(C.Structure):
    _fields_ = [("add_opcode", C.c_int32), ("mult_opcode", C.c_int32), ("xy_code", C.c_int32),
                ("z_code", C.c_int32), ("flipxy", C.c_int32)]


class _CInfo(C.Structure):
    _fields_ = [("vlen", C.c_int64), ("vdim", C.c_int64), ("nvec", C.c_int64),
                ("nvec_nonempty", C.c_int64), ("nnz", C.c_int64), ("is_hyper", C.c_int32),
                ("type_code", C.c_int32), ("method_used", C.c_int32), ("mask_applied", C.c_int32),
                ("flops", C.c_int64), ("device_ms", C.c_double), ("kernel_ms", C.c_double)]


lib.gb200_last_error.restype = C.c_char_p
lib.gb200_version.restype = C.c_char_p
lib.gb200_kernel_launches.restype = C.c_int64
lib.gb200_multiplies.restype = C.c_int64
for _name in ("gb200_init", "gb200_finalize", "gb200_upload", "gb200_upload_from_device", "gb200_dmatrix_free",
              "gb200_AxB_device", "gb200_AxB_host", "gb200_result_get_info", "gb200_result_fetch",
              "gb200_result_free", "gb200_flopcount_device", "gb200_partition_by_flops",
              "gb200_semiring_canonical", "gb200_device_count", "gb200_timer_mark",
              "gb200_timer_elapsed_ms"):
    getattr(lib, _name).restype = C.c_int
lib.gb200_host_malloc.restype = C.c_void_p
lib.gb200_host_malloc.argtypes = [C.c_size_t]
lib.gb200_host_free.restype = None
lib.gb200_host_free.argtypes = [C.c_void_p]
lib.gb200_host_trim.restype = None


def _check(code: int, where: str) -> None:
    if code != 0:
        raise GB200Error(code, where)


class _HostBlock:
    """A block from gb200_host_malloc (page-locked when large), exposed through the array interface
    so that numpy can view it; returned to the library's cache when the last view dies."""

    def __init__(self, nbytes: int):
        self.nbytes = max(int(nbytes), 1)
        self.ptr = lib.gb200_host_malloc(self.nbytes)
        if not self.ptr:
            raise MemoryError(f"gb200_host_malloc({self.nbytes})")
        self.__array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1",
                                    "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            if self.ptr:
                lib.gb200_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def host_empty(n: int, dtype) -> np.ndarray:
    """np.empty(n, dtype) on memory from gb200_host_malloc -- what a host application gets for every
    GraphBLAS array after GxB_init(mode, gb200_host_malloc, ...)."""
    dt = np.dtype(dtype)
    blk = _HostBlock(n * dt.itemsize)
    return np.asarray(blk)[: n * dt.itemsize].view(dt)


def host_array(a: np.ndarray) -> np.ndarray:
    out = host_empty(a.size, a.dtype)
    out[...] = a.reshape(-1)
    return out


@dataclass
class Semiring:
    """add monoid, multiply operator, operand type: e.g. Semiring('PLUS', 'TIMES', 'FP64')."""
    add: str
    mult: str
    xytype: str
    flipxy: bool = False

    @classmethod
    def parse(cls, name: str) -> "Semiring":
        """'PLUS_TIMES_FP64' / 'GxB_LOR_LAND_BOOL' -> Semiring."""
        parts = name.replace("GxB_", "").replace("GrB_", "").split("_")
        return cls(parts[0], parts[1], parts[2])

    @property
    def ztype(self) -> str:
        return "BOOL" if self.mult in COMPARE_OPS else self.xytype

    def c(self) -> _CSemiring:
        return _CSemiring(OPCODES[self.add], OPCODES[self.mult], TYPES[self.xytype][0],
                          TYPES[self.ztype][0], 1 if self.flipxy else 0)


@dataclass
class Matrix:
    """A host sparse matrix in the reference's CSC-agnostic layout."""
    vlen: int
    vdim: int
    p: np.ndarray
    i: np.ndarray
    x: np.ndarray
    h: Optional[np.ndarray] = None
    type: str = field(default="")

    def __post_init__(self):
        self.p = np.ascontiguousarray(self.p, dtype=np.int64)
        self.i = np.ascontiguousarray(self.i, dtype=np.int64)
        if self.h is not None:
            self.h = np.ascontiguousarray(self.h, dtype=np.int64)
        if not self.type:
            for k, (_, dt) in TYPES.items():
                if np.dtype(dt) == self.x.dtype:
                    self.type = k
        self.x = np.ascontiguousarray(self.x, dtype=TYPES[self.type][1])

    @property
    def nvec(self) -> int:
        return len(self.p) - 1

    @property
    def nnz(self) -> int:
        return int(self.p[-1])

    def c(self) -> _CMatrix:
        return _CMatrix(self.vlen, self.vdim, self.nvec, self.p.ctypes.data,
                        self.h.ctypes.data if self.h is not None else None,
                        self.i.ctypes.data if self.i.size else None,
                        self.x.ctypes.data if self.x.size else None,
                        TYPES[self.type][0], 0)

    @classmethod
    def from_scipy(cls, s, type: str = "") -> "Matrix":
        """scipy CSC -> vectors are columns; scipy CSR -> vectors are rows."""
        s.sort_indices()
        if s.format == "csc":
            vlen, vdim = s.shape
        else:
            vdim, vlen = s.shape
        return cls(vlen, vdim, s.indptr.astype(np.int64), s.indices.astype(np.int64), s.data, None,
                   type)

    def pinned(self) -> "Matrix":
        """The same matrix with its arrays in memory from gb200_host_malloc."""
        return Matrix(self.vlen, self.vdim, host_array(self.p), host_array(self.i),
                      host_array(self.x), host_array(self.h) if self.h is not None else None,
                      self.type)

    def to_hyper(self) -> "Matrix":
        """Same matrix in hypersparse form (only non-empty vectors are listed)."""
        cnt = np.diff(self.p)
        if self.h is not None:
            keep = cnt > 0
            return Matrix(self.vlen, self.vdim, np.concatenate([[0], np.cumsum(cnt[keep])]),
                          self.i, self.x, self.h[keep], self.type)
        keep = np.nonzero(cnt > 0)[0]
        return Matrix(self.vlen, self.vdim, np.concatenate([[0], np.cumsum(cnt[keep])]), self.i,
                      self.x, keep.astype(np.int64), self.type)

    def to_standard(self) -> "Matrix":
        if self.h is None:
            return self
        cnt = np.zeros(self.vdim, dtype=np.int64)
        cnt[self.h] = np.diff(self.p)
        return Matrix(self.vlen, self.vdim, np.concatenate([[0], np.cumsum(cnt)]), self.i, self.x,
                      None, self.type)


class DMatrix:
    """A matrix resident in HBM (gb200_dmatrix)."""

    def __init__(self, m: Matrix):
        self._h = C.c_void_p()
        cm = m.c()
        _check(lib.gb200_upload(C.byref(self._h), C.byref(cm)), "gb200_upload")
        self.host = m

    @classmethod
    def from_device(cls, like: Matrix, p_ptr: int, h_ptr: int, i_ptr: int, x_ptr: int) -> "DMatrix":
        """gb200_upload_from_device: the arrays of `like` (same layout, 64-bit indices) already sit
        in HBM at the given addresses, e.g. after an NCCL all-gather; `like` supplies shape and nnz."""
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        self.host = like
        cm = _CMatrix(like.vlen, like.vdim, like.nvec, p_ptr, h_ptr if like.h is not None else None,
                      i_ptr if like.nnz else None, x_ptr if like.nnz else None, TYPES[like.type][0], 0)
        _check(lib.gb200_upload_from_device(C.byref(self._h), C.byref(cm), C.c_int64(like.nnz)),
               "gb200_upload_from_device")
        return self

    def free(self):
        if self._h:
            lib.gb200_dmatrix_free(C.byref(self._h))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class Result:
    matrix: Optional[Matrix]
    info: dict


def _fetch(rh: C.c_void_p, fetch: bool, pinned: bool = False) -> Result:
    ci = _CInfo()
    _check(lib.gb200_result_get_info(rh, C.byref(ci)), "gb200_result_get_info")
    info = {k: getattr(ci, k) for k, _ in _CInfo._fields_}
    m = None
    if fetch:
        tname, dt = TYPE_BY_CODE[ci.type_code]
        empty = host_empty if pinned else np.empty
        p = empty(ci.nvec + 1, np.int64)
        h = empty(ci.nvec, np.int64) if ci.is_hyper else None
        i = empty(ci.nnz, np.int64)
        x = empty(ci.nnz, dt)
        _check(lib.gb200_result_fetch(rh, p.ctypes.data_as(C.c_void_p),
                                      h.ctypes.data_as(C.c_void_p) if h is not None else None,
                                      i.ctypes.data_as(C.c_void_p) if ci.nnz else None,
                                      x.ctypes.data_as(C.c_void_p) if ci.nnz else None),
               "gb200_result_fetch")
        m = Matrix(ci.vlen, ci.vdim, p, i, x, h, tname)
    lib.gb200_result_free(C.byref(rh))
    return Result(m, info)


def axb_device_keep(M: Optional["DMatrix"], mask_comp: bool, A: "DMatrix", B: "DMatrix",
                    semiring: Semiring, do_adotb: bool = False, method: int = METHOD_DEFAULT):
    """gb200_AxB_device, leaving T on the device: returns (handle, info).  Use fetch_into / free_result."""
    rh = C.c_void_p()
    s = semiring.c()
    _check(lib.gb200_AxB_device(C.byref(rh), M._h if M is not None else None, int(mask_comp), A._h,
                                B._h, C.byref(s), int(do_adotb), method), "gb200_AxB_device")
    ci = _CInfo()
    _check(lib.gb200_result_get_info(rh, C.byref(ci)), "gb200_result_get_info")
    return rh, {k: getattr(ci, k) for k, _ in _CInfo._fields_}


def fetch_into(rh, p_ptr: int, h_ptr: int, i_ptr: int, x_ptr: int) -> None:
    """gb200_result_fetch into caller-owned buffers given by address (host, or device memory of the
    same GPU, e.g. torch.Tensor.data_ptr())."""
    _check(lib.gb200_result_fetch(rh, C.c_void_p(p_ptr), C.c_void_p(h_ptr) if h_ptr else None,
                                  C.c_void_p(i_ptr) if i_ptr else None,
                                  C.c_void_p(x_ptr) if x_ptr else None), "gb200_result_fetch")


def free_result(rh) -> None:
    lib.gb200_result_free(C.byref(rh))


def init(device: int = -1) -> None:
    _check(lib.gb200_init(device), "gb200_init")


def axb_device(M: Optional[DMatrix], mask_comp: bool, A: DMatrix, B: DMatrix, semiring: Semiring,
               do_adotb: bool = False, method: int = METHOD_DEFAULT, fetch: bool = True,
               pinned: bool = False) -> Result:
    """C<M>=A*B (or A'*B) with operands already resident (gb200_AxB_device)."""
    rh = C.c_void_p()
    s = semiring.c()
    _check(lib.gb200_AxB_device(C.byref(rh), M._h if M is not None else None, int(mask_comp), A._h,
                                B._h, C.byref(s), int(do_adotb), method), "gb200_AxB_device")
    return _fetch(rh, fetch, pinned)


def axb_host(M: Optional[Matrix], mask_comp: bool, A: Matrix, B: Matrix, semiring: Semiring,
             do_adotb: bool = False, method: int = METHOD_DEFAULT, fetch: bool = True,
             pinned: bool = False) -> Result:
    """The GB_AxB_parallel replacement: host operands in, host T out (gb200_AxB_host + fetch).
    pinned: T's arrays come from gb200_host_malloc (as GB_create would hand out under GxB_init
    with the gb200_host_* allocator)."""
    rh = C.c_void_p()
    s = semiring.c()
    cm = M.c() if M is not None else None
    ca, cb = A.c(), (A.c() if B is A else B.c())
    _check(lib.gb200_AxB_host(C.byref(rh), C.byref(cm) if cm is not None else None, int(mask_comp),
                              C.byref(ca), C.byref(cb), C.byref(s), int(do_adotb), method),
           "gb200_AxB_host")
    return _fetch(rh, fetch, pinned)


def flopcount(M: Optional[DMatrix], A: DMatrix, B: DMatrix):
    """GB_AxB_flopcount on the device: returns (cumulative Bflops [nvec+1], total)."""
    out = np.empty(B.host.nvec + 1, dtype=np.int64)
    total = C.c_int64()
    _check(lib.gb200_flopcount_device(M._h if M is not None else None, A._h, B._h,
                                      out.ctypes.data_as(C.c_void_p), C.byref(total)),
           "gb200_flopcount_device")
    return out, total.value


def partition_by_flops(cum: np.ndarray, nparts: int) -> np.ndarray:
    cum = np.ascontiguousarray(cum, dtype=np.int64)
    bounds = np.empty(nparts + 1, dtype=np.int64)
    _check(lib.gb200_partition_by_flops(cum.ctypes.data_as(C.c_void_p), len(cum) - 1, nparts,
                                        bounds.ctypes.data_as(C.c_void_p)),
           "gb200_partition_by_flops")
    return bounds


def kernel_launches() -> int:
    return lib.gb200_kernel_launches()


def timer_mark(slot: int) -> None:
    _check(lib.gb200_timer_mark(slot), "gb200_timer_mark")


def timer_elapsed_ms(a: int, b: int) -> float:
    ms = C.c_double()
    _check(lib.gb200_timer_elapsed_ms(a, b, C.byref(ms)), "gb200_timer_elapsed_ms")
    return ms.value
