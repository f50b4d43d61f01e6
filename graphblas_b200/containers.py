"""containers.py -- the host-side containers of the binding (Matrix, Semiring) and the code tables
of include/gb_b200.h.  Pure Python + numpy + ctypes structure definitions: importing this file
does NOT load libgb_b200.so, so bench.py's `--impl reference` arm (which must not map any product
library) loads it by path, bypassing graphblas_b200/__init__.py.

Vocabulary follows the reference (SuiteSparse:GraphBLAS v2.3.3): a matrix is `vdim` sparse vectors
of length `vlen` ("CSC-agnostic", Source/Template/GB_matrix.h:193-208); `h` is the hyperlist.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

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



class _CMatrix(C.Structure):
    _fields_ = [("vlen", C.c_int64), ("vdim", C.c_int64), ("nvec", C.c_int64),
                ("p", C.c_void_p), ("h", C.c_void_p), ("i", C.c_void_p), ("x", C.c_void_p),
                ("type_code", C.c_int32), ("reserved", C.c_int32)]


class _CSemiring(C.Structure):
    _fields_ = [("add_opcode", C.c_int32), ("mult_opcode", C.c_int32), ("xy_code", C.c_int32),
                ("z_code", C.c_int32), ("flipxy", C.c_int32)]



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
        from . import host_array         # needs the CUDA library (page-locked allocator)
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
