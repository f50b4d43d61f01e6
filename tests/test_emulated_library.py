"""The WHOLE library on the host: tools/emu_library.py compiles the product sources (host orchestration of
engine_*.cu and every kernel; `<<< >>>` launches rewritten textually) against tests/emu/cuda_runtime.h
into libgb_b200_emu.so, and the GPU parity tests of tests/test_gpu_seam.py are then run through the real
ctypes binding with `graphblas_b200.lib` pointed at it (tests/emulated.py) -- against the same pinned
oracle, golden vectors and known answers as on the GPU.  A sample of them, sized for the CPU suite; any
other `-m gpu` seam test can be run the same way when developing without a GPU.

This is test infrastructure: the product package never loads the emulated library, and a green run here
is NOT a parity claim for the GPU (the memory model, the warp scheduling and the one PTX load are
emulated); it checks the engine's sequencing and the kernels' logic on every CPU run."""
import numpy as np
import pytest

import emulated
import gen
import oracle_c
import semirings
import test_gpu_seam as T


@pytest.fixture(scope="module")
def emu():
    emulated.library()          # built once (cached by a hash of the sources)
    return emulated


def test_library_identity(emu):
    with emu.swapped() as gb:
        assert gb.lib.gb200_device_count() == 1
        assert b"gb_b200" in gb.lib.gb200_version()
    import graphblas_b200 as gb
    assert gb.lib._name.endswith("libgb_b200.so")       # the swap is undone


def test_semiring_sample(emu):
    """every 60th of the 960 built-in workers: saxpy, masked saxpy, masked dot, dot against the oracle"""
    with emu.swapped() as gb:
        n = 0
        for add, mult, t in list(semirings.all_builtin())[::60]:
            dt = T.NPT[t]
            A = gb.Matrix.from_scipy(gen.er(40, 30, 260, 61, dt, lo=-3, hi=4).tocsc())
            B = gb.Matrix.from_scipy(gen.er(30, 35, 240, 62, dt, lo=-3, hi=4).tocsc())
            At = gb.Matrix.from_scipy(gen.er(30, 40, 260, 63, dt, lo=-3, hi=4).tocsc())
            M = gb.Matrix.from_scipy(gen.er(40, 35, 500, 64, np.bool_).tocsc())
            sr = gb.Semiring(add, mult, t)
            tag = f"{add}_{mult}_{t}"
            T.assert_same(oracle_c.axb(None, False, A, B, sr), gb.axb_host(None, False, A, B, sr).matrix, add, tag + " saxpy")
            T.assert_same(oracle_c.axb(M, False, A, B, sr), gb.axb_host(M, False, A, B, sr).matrix, add, tag + " masked saxpy")
            T.assert_same(oracle_c.axb(M, False, At, B, sr, True), gb.axb_host(M, False, At, B, sr, True).matrix, add, tag + " masked dot")
            T.assert_same(oracle_c.axb(None, False, At, B, sr, True), gb.axb_host(None, False, At, B, sr, True).matrix, add, tag + " dot")
            n += 1
        assert n == 16


SEAM = [
    ("flopcount", lambda: [T.test_flopcount_matches_oracle(h, m) for h in (False, True) for m in (False, True)]),
    ("golden vectors", T.test_golden_vectors_on_gpu),
    ("tri_demo known answers", T.test_tri_demo_known_answers_on_gpu),
    ("masked dot, hubs, pattern-only PLUS_TIMES_INT64", lambda: T.test_masked_dot_hubs(True, "PLUS", "TIMES", "INT64")),
    ("masked dot, hubs, valued MIN_PLUS_FP64", lambda: T.test_masked_dot_hubs(False, "MIN", "PLUS", "FP64")),
    ("masked dot, edge cases", T.test_masked_dot_edge_cases),
    ("NaN / Inf placement", T.test_nan_inf_placement),
    ("typecast INT32 x FP32 -> FP64", lambda: T.test_typecast_at_the_seam("INT32", "FP32", "PLUS", "TIMES", "FP64")),
    ("vector pull and push, PLUS_TIMES_FP64", lambda: T.test_vector_pull_and_push("PLUS", "TIMES", "FP64", False)),
    ("vector pull and push, MIN_PLUS_FP64, dense u", lambda: T.test_vector_pull_and_push("MIN", "PLUS", "FP64", True)),
    ("vector pull and push, LOR_LAND_BOOL", lambda: T.test_vector_pull_and_push("LOR", "LAND", "BOOL", False)),
]


@pytest.mark.parametrize("what,body", SEAM, ids=[s[0] for s in SEAM])
def test_seam_test_on_the_host(emu, what, body):
    with emu.swapped():
        body()


def test_owner_aligned_parts_merge_to_the_whole(emu):
    """the N-GPU cut of the masked dot (graphblas_b200/sharded.py, what bench.py --gpus N runs): every
    rank's part through the library, merged, equals the whole T and the oracle's"""
    import scipy.sparse as sps
    from graphblas_b200 import sharded
    with emu.swapped() as gb:
        S = gen.rmat_scipy(10, 8).tocsc().astype(np.int64)
        L = gb.Matrix.from_scipy(sps.tril(S, -1).tocsc())
        U = gb.Matrix.from_scipy(sps.triu(S, 1).tocsc())
        sr = gb.Semiring("PLUS", "TIMES", "INT64", True)
        ref = oracle_c.axb(L, False, U, L, sr, True)
        whole = gb.axb_host(L, False, U, L, sr, True).matrix
        T.assert_same(ref, whole, "PLUS", "whole")
        op = sharded.OwnerPartition(L, U, L, 4, hub_len=64)
        parts = [gb.axb_host(op.mask(r), False, U, L, sr, True).matrix for r in range(4)]
        merged = sharded.merge_disjoint(parts)
        T.assert_same(ref, merged, "PLUS", "merged parts")
        assert int(merged.x.sum()) == int(ref.x.sum())          # the triangle count


def test_concurrent_callers(emu):
    """SURVEY 8b, threading: the reference may be entered from many user threads (openmp_demo, pthread_demo);
    the library serialises device use behind a mutex.  Four threads multiply different operands at once
    (ctypes releases the GIL around the calls) and each gets its own right answer."""
    import threading
    with emu.swapped() as gb:
        jobs = []
        for k in range(4):
            A = gb.Matrix.from_scipy(gen.er(120, 90, 1500, 700 + k, np.int64, lo=1, hi=6).tocsc())
            B = gb.Matrix.from_scipy(gen.er(90, 100, 1400, 710 + k, np.int64, lo=1, hi=6).tocsc())
            At = gb.Matrix.from_scipy(gen.er(90, 120, 1500, 720 + k, np.int64, lo=1, hi=6).tocsc())
            M = gb.Matrix.from_scipy(gen.er(120, 100, 2500, 730 + k, np.bool_).tocsc())
            jobs.append((A, B, At, M))
        sr = gb.Semiring("PLUS", "TIMES", "INT64")
        want = [(oracle_c.axb(M, False, A, B, sr), oracle_c.axb(M, False, At, B, sr, True)) for A, B, At, M in jobs]
        got, errs = [None] * 4, []

        def work(k):
            try:
                A, B, At, M = jobs[k]
                for _ in range(2):
                    got[k] = (gb.axb_host(M, False, A, B, sr).matrix, gb.axb_host(M, False, At, B, sr, True).matrix)
            except Exception as e:      # surfaced below, in the main thread
                errs.append(repr(e))
        th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
        for t in th:
            t.start()
        for t in th:
            t.join(600)
        assert not errs, errs
        for k in range(4):
            T.assert_same(want[k][0], got[k][0], "PLUS", f"thread {k} masked saxpy")
            T.assert_same(want[k][1], got[k][1], "PLUS", f"thread {k} masked dot")


def test_transpose_and_accum_mask_against_the_oracle(emu):
    """rows f2 / f1 through the C ABI: gb200_transpose_host and gb200_accum_mask_host against the pinned
    restatements (a sample of tests/test_gpu_seam.py)"""
    with emu.swapped():
        T.test_transpose_matches_oracle(True, "INT16")
        T.test_transpose_matches_oracle(False, None)
        T.test_accum_mask_matches_oracle(True, True, True)
        T.test_accum_mask_matches_oracle(False, False, False)
        T.test_accum_mask_edge_cases()
