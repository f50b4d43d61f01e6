"""gen.py -- deterministic synthetic inputs shared by the tests and bench.py.

ER: uniformly random pattern (the shape of reference Demo/Source/random_matrix.c inputs).
RMAT: Graph500 Kronecker generator (a,b,c,d) = (0.57,0.19,0.19,0.05), counter-based splitmix64 RNG,
hashed vertex relabelling, self-loops dropped, symmetrised, de-duplicated.  Written
with torch so the same code runs on the CPU (tests, small scales) and on the GPU (bench, scale 22).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

_M64 = (1 << 64) - 1


def er(nrows: int, ncols: int, nnz: int, seed: int, dtype=np.float64, lo=-4, hi=5) -> sp.csr_matrix:
    rng = np.random.default_rng(seed)
    i = rng.integers(0, nrows, nnz)
    j = rng.integers(0, ncols, nnz)
    key = np.unique(i.astype(np.int64) * ncols + j)
    i, j = key // ncols, key % ncols
    dt = np.dtype(dtype)
    if dt == np.bool_:
        x = rng.integers(0, 2, len(key)).astype(np.bool_)
    elif dt.kind in "iu":
        l = max(lo, 0) if dt.kind == "u" else lo
        x = rng.integers(l, hi, len(key)).astype(dt)
    else:
        x = (rng.random(len(key)) * 2 - 0.5).astype(dt)
    m = sp.csr_matrix((x, (i, j)), shape=(nrows, ncols))
    m.sort_indices()
    return m


def _splitmix64_t(x):
    """splitmix64 finaliser on a torch int64 tensor (wrap-around arithmetic, logical shifts)."""
    import torch

    def lsr(v, s):
        return (v >> s) & ((1 << (64 - s)) - 1)

    def c(v):        # python int -> wrapped int64 constant
        v &= _M64
        return v - (1 << 64) if v >= (1 << 63) else v

    x = x + c(0x9E3779B97F4A7C15)
    x = (x ^ lsr(x, 30)) * c(0xBF58476D1CE4E5B9)
    x = (x ^ lsr(x, 27)) * c(0x94D049BB133111EB)
    x = x ^ lsr(x, 31)
    return x


def _u01(h):
    """int64 hash -> float64 in [0,1)"""
    import torch
    return ((h >> 11) & ((1 << 53) - 1)).to(torch.float64) * (1.0 / (1 << 53))


def rmat_edges(scale: int, edgefactor: int = 16, seed: int = 42, device: str = "cpu"):
    """Symmetric, loop-free, de-duplicated RMAT graph: returns (n, rows, cols) with (rows, cols)
    sorted by row then column (torch int64 tensors on `device`)."""
    import torch
    n = 1 << scale
    m = edgefactor * n
    eid = torch.arange(m, dtype=torch.int64, device=device)
    ii = torch.zeros(m, dtype=torch.int64, device=device)
    jj = torch.zeros(m, dtype=torch.int64, device=device)
    a, b, c = 0.57, 0.19, 0.19
    for level in range(scale):
        r = _u01(_splitmix64_t(_splitmix64_t(eid * 64 + level) ^ seed))
        ibit = (r >= a + b).to(torch.int64)
        jbit = (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)
        ii = ii * 2 + ibit
        jj = jj * 2 + jbit
    # hashed relabelling: vertex v -> rank of hash(v)
    hv = _splitmix64_t(torch.arange(n, dtype=torch.int64, device=device) ^ (seed + 1))
    perm = torch.empty(n, dtype=torch.int64, device=device)
    perm[torch.argsort(hv)] = torch.arange(n, dtype=torch.int64, device=device)
    ii, jj = perm[ii], perm[jj]
    keep = ii != jj
    ii, jj = ii[keep], jj[keep]
    key = torch.cat([ii * n + jj, jj * n + ii])
    key = torch.unique(key)                     # sorted
    return n, key // n, key % n


def rmat_weights(rows, cols, seed: int = 44):
    """Symmetric weights in (0,1] from a hash of the unordered vertex pair."""
    import torch
    lo, hi = torch.minimum(rows, cols), torch.maximum(rows, cols)
    h = _splitmix64_t(_splitmix64_t(lo * 0x1000003 + hi) ^ seed)
    return (((h >> 11) & ((1 << 53) - 1)).to(torch.float64) + 1.0) * (1.0 / (1 << 53))


def csr_from_sorted(n: int, rows, cols):
    """row pointer array from sorted (rows, cols) torch tensors"""
    import torch
    counts = torch.bincount(rows, minlength=n)
    p = torch.zeros(n + 1, dtype=torch.int64, device=rows.device)
    p[1:] = torch.cumsum(counts, 0)
    return p


def rmat_scipy(scale: int, edgefactor: int = 16, seed: int = 42, weighted: bool = False,
               dtype=np.float64) -> sp.csr_matrix:
    n, r, c = rmat_edges(scale, edgefactor, seed)
    if weighted:
        x = rmat_weights(r, c).numpy().astype(dtype)
    else:
        x = np.ones(len(r), dtype=dtype)
    m = sp.csr_matrix((x, (r.numpy(), c.numpy())), shape=(n, n))
    m.sort_indices()
    return m
