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

from .containers import (TYPES, TYPE_BY_CODE, OPCODES, COMPARE_OPS, METHOD_DEFAULT,  # noqa: F401
                         METHOD_GUSTAVSON, METHOD_HEAP, METHOD_DOT, STATUS, _CMatrix, _CSemiring,
                         Semiring, Matrix)


class GB200Error(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib.gb200_last_error().decode(errors="replace")
        super().__init__(f"{where}: GB200_{STATUS.get(code, code)}: {msg}")


class _CInfo(C.Structure):
    _fields_ = [("vlen", C.c_int64), ("vdim", C.c_int64), ("nvec", C.c_int64),
                ("nvec_nonempty", C.c_int64), ("nnz", C.c_int64), ("is_hyper", C.c_int32),
                ("type_code", C.c_int32), ("method_used", C.c_int32), ("mask_applied", C.c_int32),
                ("flops", C.c_int64), ("device_ms", C.c_double), ("kernel_ms", C.c_double)]


# every entry point of include/gb_b200.h with its full signature (without argtypes ctypes would pass
# Python ints as 32-bit C ints where the ABI takes int64_t)
_VP, _I, _I64 = C.c_void_p, C.c_int, C.c_int64
_SIGS = {
    "gb200_last_error": (C.c_char_p, []),
    "gb200_version": (C.c_char_p, []),
    "gb200_kernel_launches": (_I64, []),
    "gb200_multiplies": (_I64, []),
    "gb200_device_count": (_I, []),
    "gb200_init": (_I, [_I]),
    "gb200_finalize": (_I, []),
    "gb200_semiring_canonical": (_I, [_VP]),
    "gb200_upload": (_I, [_VP, _VP]),
    "gb200_upload_from_device": (_I, [_VP, _VP, _I64]),
    "gb200_dmatrix_free": (_I, [_VP]),
    "gb200_AxB_device": (_I, [_VP, _VP, _I, _VP, _VP, _VP, _I, _I]),
    "gb200_AxB_host": (_I, [_VP, _VP, _I, _VP, _VP, _VP, _I, _I]),
    "gb200_result_get_info": (_I, [_VP, _VP]),
    "gb200_result_fetch": (_I, [_VP, _VP, _VP, _VP, _VP]),
    "gb200_result_free": (_I, [_VP]),
    "gb200_flopcount_device": (_I, [_VP, _VP, _VP, _VP, _VP]),
    "gb200_partition_by_flops": (_I, [_VP, _I64, _I, _VP]),
    "gb200_timer_mark": (_I, [_I]),
    "gb200_timer_elapsed_ms": (_I, [_I, _I, _VP]),
    "gb200_host_malloc": (_VP, [C.c_size_t]),
    "gb200_host_calloc": (_VP, [C.c_size_t, C.c_size_t]),
    "gb200_host_realloc": (_VP, [_VP, C.c_size_t]),
    "gb200_host_free": (None, [_VP]),
    "gb200_host_trim": (None, []),
    "gb200_device_trim": (None, []),
    "gb200_reduce_device": (_I, [_VP, _I, _VP]),
    "gb200_reduce_host": (_I, [_VP, _I, _VP]),
    "gb200_result_adopt": (_I, [_VP, _VP]),
    "gb200_result_reduce": (_I, [_VP, _I, _VP]),
    "gb200_select_device": (_I, [_VP, _VP, _I, _I64]),
    "gb200_select_host": (_I, [_VP, _VP, _I, _I64]),
    "gb200_accum_mask_device": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _I]),
    "gb200_accum_mask_host": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _I, _I]),
    "gb200_assign_scalar_device": (_I, [_VP, _VP, _VP, _I, _I, _I, _VP, _I, _I]),
    "gb200_assign_scalar_host": (_I, [_VP, _VP, _VP, _I, _I, _I, _VP, _I, _I]),
    "gb200_transpose_device": (_I, [_VP, _VP, _I, _I, C.c_double]),
    "gb200_transpose_host": (_I, [_VP, _VP, _I, _I, C.c_double]),
    "gb200_peerbuf_create": (_I, [_VP, _I64, _I, _I, _I]),
    "gb200_peerbuf_handle": (_I, [_VP, _VP]),
    "gb200_peerbuf_connect": (_I, [_VP, _VP]),
    "gb200_peerbuf_publish": (_I, [_VP, _VP]),
    "gb200_peerbuf_wait": (_I, [_VP]),
    "gb200_peerbuf_view": (_I, [_VP, _VP, _VP, _VP]),
    "gb200_peerbuf_read": (_I, [_VP, _VP, _VP]),
    "gb200_peerbuf_free": (_I, [_VP]),
    "gb200_cache_enable": (None, [_I]),
    "gb200_cache_enabled": (_I, []),
    "gb200_cache_invalidate": (None, [_VP]),
    "gb200_cache_clear": (None, []),
    "gb200_cache_stats": (None, [_VP, _VP, _VP, _VP]),
}
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.restype, _fn.argtypes = _res, _args


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
    try:
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
        return Result(m, info)
    finally:
        lib.gb200_result_free(C.byref(rh))      # also when the fetch failed


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


SELECT_OPS = {"TRIL": 0, "TRIU": 1, "DIAG": 2, "OFFDIAG": 3, "NONZERO": 4}


def select_host(A: Matrix, op: str, k: int = 0, pinned: bool = False) -> Result:
    """GxB_select with a built-in operator on the device (gb200_select_host + fetch); `A` in the
    CSC-agnostic layout: TRIL keeps (vector - index) <= k, and so on."""
    rh = C.c_void_p()
    ca = A.c()
    _check(lib.gb200_select_host(C.byref(rh), C.byref(ca), SELECT_OPS[op], k), "gb200_select_host")
    return _fetch(rh, True, pinned)


def transpose_host(A: Matrix, ctype: Optional[str] = None, hyper: bool = False, hyper_ratio: float = -1.0,
                   pinned: bool = False) -> Result:
    """C = (ctype) A' on the device (gb200_transpose_host + fetch), reference Source/GB_transpose.c: in the
    CSC-agnostic layout entry i of vector j becomes entry j of vector i; hyper: C lists only its
    non-empty vectors; hyper_ratio >= 0: that form is then conformed as GB_to_hyper_conform would."""
    rh = C.c_void_p()
    ca = A.c()
    code = TYPES[ctype][0] if ctype is not None else ca.type_code
    _check(lib.gb200_transpose_host(C.byref(rh), C.byref(ca), code, int(hyper), float(hyper_ratio)),
           "gb200_transpose_host")
    return _fetch(rh, True, pinned)


def accum_mask_host(Cm: Matrix, T: Matrix, M: Optional[Matrix] = None, mask_comp: bool = False,
                    replace: bool = False, accum: Optional[tuple] = None, hyper: bool = False,
                    pinned: bool = False) -> Result:
    """C<M> = accum (C,T) on the device (gb200_accum_mask_host + fetch), reference Source/GB_accum_mask.c:
    accum = (operator name, type name of its inputs) or None; returns the new C"""
    rh = C.c_void_p()
    cc, ct = Cm.c(), T.c()
    cm = M.c() if M is not None else None
    op, xy = (OPCODES[accum[0]], TYPES[accum[1]][0]) if accum is not None else (0, 0)
    _check(lib.gb200_accum_mask_host(C.byref(rh), C.byref(cc), C.byref(ct),
                                     C.byref(cm) if cm is not None else None, int(mask_comp), int(replace),
                                     op, xy, int(hyper)), "gb200_accum_mask_host")
    return _fetch(rh, True, pinned)


def assign_scalar_host(Cm: Matrix, M: Matrix, scalar, scalar_type: str, replace: bool = False,
                       accum: Optional[tuple] = None, hyper: bool = False, pinned: bool = False) -> Result:
    """C<M> = accum (C, scalar) over all of C on the device (gb200_assign_scalar_host + fetch), reference
    GrB_assign with GrB_ALL and a scalar (Source/GB_assign_scalar.c); returns the new C"""
    rh = C.c_void_p()
    cc, cm = Cm.c(), M.c()
    x = np.array([scalar], dtype=TYPES[scalar_type][1])
    op, xy = (OPCODES[accum[0]], TYPES[accum[1]][0]) if accum is not None else (0, 0)
    _check(lib.gb200_assign_scalar_host(C.byref(rh), C.byref(cc), C.byref(cm), int(replace), op, xy,
                                        x.ctypes.data_as(C.c_void_p), TYPES[scalar_type][0], int(hyper)),
           "gb200_assign_scalar_host")
    return _fetch(rh, True, pinned)


def accum_mask_device(Cm: DMatrix, T: DMatrix, M: Optional[DMatrix] = None, mask_comp: bool = False,
                      replace: bool = False, accum: Optional[tuple] = None, hyper: bool = False,
                      fetch: bool = True, pinned: bool = False) -> Result:
    """the same with C, T and M already resident in HBM (gb200_accum_mask_device); fetch=False: timing only"""
    rh = C.c_void_p()
    op, xy = (OPCODES[accum[0]], TYPES[accum[1]][0]) if accum is not None else (0, 0)
    _check(lib.gb200_accum_mask_device(C.byref(rh), Cm._h, T._h, M._h if M is not None else None,
                                       int(mask_comp), int(replace), op, xy, int(hyper)),
           "gb200_accum_mask_device")
    return _fetch(rh, fetch, pinned)


def transpose_device(A: DMatrix, ctype: Optional[str] = None, hyper: bool = False, hyper_ratio: float = -1.0,
                     fetch: bool = True, pinned: bool = False) -> Result:
    """the same with A already resident in HBM (gb200_transpose_device); fetch=False: timing only"""
    rh = C.c_void_p()
    code = TYPES[ctype][0] if ctype is not None else TYPES[A.host.type][0]
    _check(lib.gb200_transpose_device(C.byref(rh), A._h, code, int(hyper), float(hyper_ratio)),
           "gb200_transpose_device")
    return _fetch(rh, fetch, pinned)


def result_reduce(rh, type_: str, add: str = "PLUS"):
    """monoid reduction of the values of a result that is still on the device (gb200_result_reduce)"""
    out = np.zeros(1, dtype=TYPES[type_][1])
    _check(lib.gb200_result_reduce(rh, OPCODES[add], out.ctypes.data_as(C.c_void_p)), "gb200_result_reduce")
    return out[0]


def reduce_host(A: Matrix, add: str):
    """GrB_reduce of a matrix to a scalar over a built-in monoid, on the device (gb200_reduce_host)"""
    out = np.zeros(1, dtype=TYPES[A.type][1])
    ca = A.c()
    _check(lib.gb200_reduce_host(C.byref(ca), OPCODES[add], out.ctypes.data_as(C.c_void_p)),
           "gb200_reduce_host")
    return out[0]


class PeerBuf:
    """gb200_peerbuf: a dense copy of an n-vector on every GPU of the box, written by all ranks over
    NVLink peer memory.  `allgather_bytes(b: bytes) -> list[bytes]` exchanges the 64-byte IPC handles
    (e.g. torch.distributed.all_gather_object)."""

    def __init__(self, n: int, type_: str, rank: int, world: int, allgather_bytes):
        self._h = C.c_void_p()
        self.n, self.type, self.rank, self.world = n, type_, rank, world
        _check(lib.gb200_peerbuf_create(C.byref(self._h), n, TYPES[type_][0], rank, world),
               "gb200_peerbuf_create")
        mine = C.create_string_buffer(64)
        _check(lib.gb200_peerbuf_handle(self._h, mine), "gb200_peerbuf_handle")
        everyone = allgather_bytes(mine.raw)
        blob = C.create_string_buffer(b"".join(everyone), 64 * world)
        _check(lib.gb200_peerbuf_connect(self._h, blob), "gb200_peerbuf_connect")

    def publish(self, result_handle) -> None:
        _check(lib.gb200_peerbuf_publish(self._h, result_handle), "gb200_peerbuf_publish")

    def wait(self) -> None:
        _check(lib.gb200_peerbuf_wait(self._h), "gb200_peerbuf_wait")

    def view(self):
        """(values device pointer, presence device pointer, tag) of the current epoch; synchronises"""
        v, p, t = C.c_void_p(), C.c_void_p(), C.c_int()
        _check(lib.gb200_peerbuf_view(self._h, C.byref(v), C.byref(p), C.byref(t)), "gb200_peerbuf_view")
        return v.value, p.value, t.value

    def read(self):
        """(values [n], present [n] bool) of the current epoch on the host"""
        vals = np.empty(self.n, dtype=TYPES[self.type][1])
        pres = np.empty(self.n, dtype=np.uint8)
        _check(lib.gb200_peerbuf_read(self._h, vals.ctypes.data_as(C.c_void_p),
                                      pres.ctypes.data_as(C.c_void_p)), "gb200_peerbuf_read")
        return vals, pres != 0

    def free(self):
        if self._h:
            lib.gb200_peerbuf_free(C.byref(self._h))
            self._h = C.c_void_p()


def cache_enable(on: bool) -> None:
    """operand residency across axb_host calls (gb200_cache_enable)"""
    lib.gb200_cache_enable(1 if on else 0)


def cache_stats() -> dict:
    v = [C.c_int64() for _ in range(4)]
    lib.gb200_cache_stats(*[C.byref(x) for x in v])
    return dict(zip(("hits", "misses", "invalidations", "resident_bytes"), [x.value for x in v]))


def kernel_launches() -> int:
    return lib.gb200_kernel_launches()


def timer_mark(slot: int) -> None:
    _check(lib.gb200_timer_mark(slot), "gb200_timer_mark")


def timer_elapsed_ms(a: int, b: int) -> float:
    ms = C.c_double()
    _check(lib.gb200_timer_elapsed_ms(a, b, C.byref(ms)), "gb200_timer_elapsed_ms")
    return ms.value
