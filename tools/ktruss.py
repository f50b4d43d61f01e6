"""tools/ktruss.py -- the k-truss loop of the reference (Extras/ktruss/ktruss_graphblas.c:96-136) as a
caller of the path on one GPU: every iteration is the masked multiply C<C> = C*C over PLUS_LAND_INT64
(the reference's GrB_mxm call at :103: no transpose, so GB_AxB_meta hands GB_AxB_parallel a masked saxpy
with VALUED operands -- the supports of the previous iteration), followed by the support filter
C = C .* (C >= k-2) (a user-defined GxB_SelectOp in the reference, :80 and :115: a host function pointer,
so it stays on the host here as well).  SURVEY.md 8f row f4.

Prints one JSON line: per-iteration device time of the multiply, multiply-adds, entries kept, the number
of steps, and -- with --check-scale -- a comparison of EVERY iteration's T at a smaller scale against
supports computed with scipy on the host ((C*C) .* C over the patterns: PLUS_LAND counts the common
neighbours).  Not the driver's bench contract (bench.py); a measurement of a caller of the path.

    python tools/ktruss.py --scale 18 --k 4 --check-scale 13 --out gpurun_out/kt/ktruss_s18.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def ktruss(gb, A: sp.csr_matrix, k: int, check: bool, max_steps: int = 100):
    n = A.shape[0]
    sr = gb.Semiring("PLUS", "LAND", "INT64")
    C = A.copy().tocsr().astype(np.int64)
    C.data[:] = 1
    C.sort_indices()
    steps, last = [], C.nnz
    for it in range(1, max_steps + 1):
        m = gb.Matrix(n, n, C.indptr.astype(np.int64), C.indices.astype(np.int64), C.data.astype(np.int64), None,
                      "INT64")
        d = gb.DMatrix(m)
        t0 = time.perf_counter()
        r = gb.axb_device(d, False, d, d, sr, False, fetch=True)
        wall = time.perf_counter() - t0
        d.free()
        T = r.matrix
        row = {"step": it, "nnz_C": int(C.nnz), "device_ms": r.info["device_ms"], "kernel_ms": r.info["kernel_ms"],
               "madds": int(r.info["flops"]), "nnz_T": int(T.nnz), "mask_applied": int(r.info["mask_applied"]),
               "wall_ms_with_fetch": wall * 1e3}
        if check:
            P = C.copy()
            P.data[:] = 1
            S = (P @ P).multiply(P).tocsr()         # supports: common neighbours of every edge
            S.eliminate_zeros()
            S.sort_indices()
            same = (np.array_equal(S.indptr, T.p) and np.array_equal(S.indices, T.i)
                    and np.array_equal(S.data.astype(np.int64), T.x))
            row["identical_to_host_supports"] = bool(same)
            if not same:
                raise SystemExit(f"k-truss step {it}: T differs from the host supports")
        keep = T.x >= (k - 2)
        rows = np.repeat(np.arange(n), np.diff(T.p))[keep]
        C = sp.csr_matrix((T.x[keep], (rows, T.i[keep])), shape=(n, n))
        C.sort_indices()
        row["kept"] = int(C.nnz)
        steps.append(row)
        if C.nnz == last:
            break
        last = C.nnz
    return C, steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=18)
    ap.add_argument("--ef", type=int, default=16)
    ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--check-scale", type=int, default=13, help="0: no host check")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import gen
    import graphblas_b200 as gb
    gb.init(0)
    out = {"workload": f"k-truss (k={args.k}) of RMAT scale {args.scale} edgefactor {args.ef}: C<C>=C*C PLUS_LAND_INT64 "
                       "(masked saxpy, valued operands) + support filter per step",
           "reference": "Extras/ktruss/ktruss_graphblas.c:96-136"}
    if args.check_scale > 0:
        A = gen.rmat_scipy(args.check_scale, args.ef, dtype=np.int64)
        C, steps = ktruss(gb, A, args.k, True)
        out["check"] = {"scale": args.check_scale, "steps": len(steps), "edges_in_truss": int(C.nnz),
                        "identical_every_step": all(s["identical_to_host_supports"] for s in steps)}
    A = gen.rmat_scipy(args.scale, args.ef, dtype=np.int64)
    ktruss(gb, A, args.k, False, max_steps=1)           # warm the allocator
    C, steps = ktruss(gb, A, args.k, False)
    madds = sum(s["madds"] for s in steps)
    dev = sum(s["device_ms"] for s in steps)
    out.update(n=int(A.shape[0]), nnz_A=int(A.nnz), nsteps=len(steps), edges_in_truss=int(C.nnz),
               madds_total=int(madds), device_ms_total=dev, gflops=2.0 * madds / dev / 1e6 if dev > 0 else None,
               steps=steps)
    line = json.dumps(out)
    print(line)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
