"""Size-independent properties at the sizes of BASELINE.json's configurations (RMAT scale 20 / 22,
edge factor 16), where the oracle would take minutes: the masked dot against the masked saxpy (the
reference's own cross-check, Demo/Program/tri_demo.c:151-155) and against its own slices, one SSSP
relaxation against a numpy segmented minimum, the BFS level loop against a numpy BFS."""
import argparse
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import graphblas_b200 as gb

pytestmark = pytest.mark.gpu


def _workload(name, scale):
    import bench
    import torch
    args = argparse.Namespace(workload=name, scale=scale, ef=16, bfs_dir="push")
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    return bench.make_workload(gb, args, dev)


def _same(a: gb.Matrix, b: gb.Matrix, what):
    assert np.array_equal(a.p, b.p), what + ": pointers"
    assert np.array_equal(a.i, b.i), what + ": pattern"
    assert np.array_equal(a.x, b.x), what + ": values"


def test_tricount_scale20_dot_equals_saxpy_and_slices():
    """config 2: C<L> = L*U' (dot) == C<L> = L*L (masked saxpy), entry for entry; the triangle count
    is the number of matched index pairs; the multiply over 3 slices of the mask's vectors
    concatenates to the same T (the N-GPU partition)"""
    w = _workload("tri", 20)
    L, U, sr = w["M"], w["A"], w["semiring"]
    dL, dU = gb.DMatrix(L), gb.DMatrix(U)
    dot = gb.axb_device(dL, False, dU, dL, sr, True)
    outer = gb.axb_device(dL, False, dL, dL, sr, False)
    assert dot.info["method_used"] == gb.METHOD_DOT and dot.info["mask_applied"] == 1
    assert outer.info["mask_applied"] == 1
    _same(dot.matrix, outer.matrix, "dot vs masked saxpy")
    ntri = int(dot.matrix.x.sum())
    assert ntri == dot.info["flops"] and ntri > 0
    import bench
    cuts = [0, L.nvec // 3, 2 * L.nvec // 3, L.nvec]
    parts = [gb.axb_host(bench.slice_vectors(gb, L, cuts[k], cuts[k + 1]), False, U, L, sr, True).matrix
             for k in range(3)]
    assert sum(int(t.x.sum()) for t in parts) == ntri
    p = np.sum([t.p for t in parts], axis=0)
    assert np.array_equal(p, dot.matrix.p)
    assert np.array_equal(np.concatenate([t.i for t in parts]), dot.matrix.i)
    assert np.array_equal(np.concatenate([t.x for t in parts]), dot.matrix.x)


def test_sssp_scale22_relaxation_is_bit_exact():
    """config 4: d' = A min.+ d with a realistic iterate d (three relaxations from the source):
    every entry equals the numpy segmented minimum of w + d[col] bit for bit (one rounding per term,
    MIN is order-independent), and the pattern is exactly the non-empty rows"""
    w = _workload("sssp", 22)
    A, sr = w["A"], w["semiring"]
    n = A.vlen
    dA = gb.DMatrix(A)
    d = w["B"].x.copy()
    lens = np.diff(A.p)
    has = lens > 0                                       # d is dense: every stored entry is a product
    for _ in range(3):
        t = A.x + d[A.i]
        ref = np.minimum.reduceat(np.concatenate([t, [np.inf]]), np.minimum(A.p[:-1], len(t)))
        dv = gb.Matrix(n, 1, np.array([0, n]), np.arange(n), d, None, "FP64")
        got = gb.axb_device(None, False, dA, gb.DMatrix(dv), sr, True).matrix
        assert np.array_equal(got.i, np.nonzero(has)[0])
        assert np.array_equal(got.x, ref[has])
        d = np.minimum(d, np.where(has, ref, np.inf))    # accum = GrB_MIN_FP64 (applied by GB_mxm)


def test_bfs_scale20_levels_match_numpy():
    """config 3: the bfs5m level loop through q<!v> = q*A; the frontier of every level equals the
    numpy BFS frontier, for the push (saxpy) and the pull (dot) direction"""
    import bench
    w = _workload("bfs", 20)
    A = w["A"]
    dA = gb.DMatrix(A)
    host = bench.host_bfs_levels(A, w["bfs_source"])
    for pull in (False, True):
        w["do_adotb"] = pull
        got = bench.bfs_levels(gb, w, dA)
        assert len(got) == len(host)
        for (q, v), (hq, hv) in zip(got, host):
            assert np.array_equal(q.i, hq) and np.array_equal(v.i, hv)
