"""tools/accum_mask_bench.py -- C<M> = accum (C,T) on one GPU at the headline scale (the
step GB_mxm runs right after the multiply, reference Source/GB_accum_mask.c:130-328).

Operands resident in HBM, built from the RMAT graph of bench.py (scale 22, edge factor 16; L = tril (A,-1),
U = triu (A,1), INT64 ones); timed: gb200_accum_mask_device (CUDA events of the library, best and median of
--reps after 2 warm-ups).  Three cases:
  same     C<L> = C + T with C = T = L: every entry is accumulated; the result is L with every value 2
  union    C = C + T with C = L, T = U, no mask: disjoint patterns; the result is A with every value 1
  vector   d = min (d, t), n-by-1, every entry present (the accumulate step of an SSSP relaxation)
Each result is checked outside the timed region against what the case must give, bit for bit, and at
--check-scale the first two against the oracle (whole T, form included).  Roofline: algorithmic bytes = C, T,
M and the result each touched once in the API layout (8-byte pointers and indices).  Prints one JSON line; not
the driver's bench contract (bench.py), a measurement of a neighbour of the path.

    python tools/accum_mask_bench.py --scale 22 --check-scale 14 --out gpurun_out/am/accum_mask_s22.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def api_bytes(*mats):
    return int(sum(8 * len(m.p) + m.nnz * (8 + m.x.dtype.itemsize) for m in mats if m is not None))


def cpu_reference(operands, scale, gpu_cases):
    """the compiled reference's own GB_accum_mask (oracle/_ref, one thread: GB_add and GB_mask are sequential
    merges) on the cases same and union at `scale`, pending work assembled inside the timed region"""
    import ctypes as C
    import time
    import grbref
    if not grbref.available():
        return {"unavailable": "oracle/_ref not built"}
    G = grbref.GraphBLAS.get(with_shim=False)
    n, L, U, A = operands(scale)
    fn = G.lib.GB_accum_mask
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_bool, C.c_bool, C.c_void_p]
    imp = lambda m: G.matrix_import("CSC", "INT64", n, n, m.p, m.i, m.x)
    out = {"cores": 1, "kind": "reference", "unit": "ms",
           "sample": f"GB_accum_mask at scale {scale} (nnz(L) = {L.nnz}), one run per case"}
    for name, Cm, T, M in (("same", L, L, L), ("union", L, U, None)):
        c, t = imp(Cm), imp(T)
        m = imp(M) if M is not None else None
        th = C.c_void_p(t.value)
        t0 = time.perf_counter()
        G.ok(fn(c, m, None, G.obj("GrB_PLUS_INT64"), C.byref(th), False, False, None), "GB_accum_mask")
        nv = G.matrix_nvals(c)
        dt = time.perf_counter() - t0
        for h in (c, m):
            if h is not None:
                G.matrix_free(h)
        ents = Cm.nnz + T.nnz
        gpu = gpu_cases.get(name, {})
        out[name] = {"value": dt * 1e3, "entries_out": int(nv), "ns_per_entry_in": dt * 1e9 / ents,
                     "gpu_ns_per_entry_in": (gpu["best_ms"] * 1e6 / gpu["entries_in"]) if gpu else None}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--ef", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check-scale", type=int, default=14)
    ap.add_argument("--out", default="")
    ap.add_argument("--cpu-scale", type=int, default=20, help="the reference's own GB_accum_mask on the host, "
                    "cases same and union at this scale (0: skip)")
    args = ap.parse_args()
    import torch
    import graphblas_b200 as gb
    import bench
    import oracle_c
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"      # cpu: only under tools/emu_*.py
    gb.init(0)
    peak_gbs, peak_source = bench.measured_peak()
    line = {"metric": "GB_accum_mask on the device", "scale": args.scale, "edgefactor": args.ef, "cases": {}}
    acc = ("PLUS", "INT64")

    def operands(scale):
        g = bench.build_rmat(scale, args.ef, dev)
        (n, Lp, Li, Lx), (_, Up, Ui, Ux) = bench.tri_operands(g)
        A = gb.Matrix(n, n, g["p"].cpu().numpy(), g["cols"].cpu().numpy(),
                      np.ones(len(g["cols"]), dtype=np.int64), None, "INT64")
        return n, gb.Matrix(n, n, Lp, Li, Lx, None, "INT64"), gb.Matrix(n, n, Up, Ui, Ux, None, "INT64"), A

    # parity at a size the oracle runs at
    n, L, U, A = operands(args.check_scale)
    for Cm, T, M, comp, rep, a in ((L, L, L, False, False, acc), (L, U, None, False, False, acc),
                                   (L, A, U, True, True, None), (A, L, U, False, False, ("MIN", "INT64"))):
        ref = oracle_c.accum_mask(Cm, T, M, comp, rep, a, False)
        got = gb.accum_mask_host(Cm, T, M, comp, rep, a, False).matrix
        if not (np.array_equal(ref.p, got.p) and np.array_equal(ref.i, got.i) and np.array_equal(ref.x, got.x)
                and ref.type == got.type):
            raise SystemExit(f"accum_mask differs from the oracle at scale {args.check_scale}")
    line["parity"] = {"oracle_scale": args.check_scale, "cases": 4, "identical": True}

    def timed(name, Cm, T, M, a, check):
        dC, dT = gb.DMatrix(Cm), (None if T is Cm else gb.DMatrix(T))
        dM = None if M is None else (dC if M is Cm else gb.DMatrix(M))
        ms = []
        for r in range(args.reps + 2):
            res = gb.accum_mask_device(dC, dT or dC, dM, False, False, a, False, fetch=False)
            if r >= 2:
                ms.append(res.info["device_ms"])
        R = gb.accum_mask_device(dC, dT or dC, dM, False, False, a, False, fetch=True).matrix
        for d in (dC, dT, dM):
            if d is not None and d is not dC:
                d.free()
        dC.free()
        if not check(R):
            raise SystemExit(f"accum_mask case {name}: the result is not what the case must give")
        algo = api_bytes(Cm, T, M, R)
        best = min(ms)
        line["cases"][name] = {"entries_in": int(Cm.nnz + T.nnz), "entries_out": int(R.nnz), "device_ms": ms,
                               "best_ms": best, "median_ms": float(np.median(ms)), "result_checked": True,
                               "roofline": {"bound": "hbm", "achieved": algo / best / 1e6, "peak": peak_gbs,
                                            "unit": "GB/s", "frac": algo / best / 1e6 / peak_gbs,
                                            "algorithmic_bytes": algo, "peak_source": peak_source}}

    n, L, U, A = operands(args.scale)
    line.update(n=n, nnz_L=int(L.nnz), nnz_A=int(A.nnz))
    timed("same", L, L, L, acc,
          lambda R: np.array_equal(R.p, L.p) and np.array_equal(R.i, L.i) and bool((R.x == 2).all()))
    timed("union", L, U, None, acc,
          lambda R: np.array_equal(R.p, A.p) and np.array_equal(R.i, A.i) and bool((R.x == 1).all()))
    del L, U, A
    rng = np.random.default_rng(7)
    d = rng.random(n)
    t = rng.random(n)
    idx = np.arange(n, dtype=np.int64)
    p1 = np.array([0, n], dtype=np.int64)
    dv, tv = gb.Matrix(n, 1, p1, idx, d, None, "FP64"), gb.Matrix(n, 1, p1, idx, t, None, "FP64")
    timed("vector", dv, tv, None, ("MIN", "FP64"),
          lambda R: np.array_equal(R.i, idx) and np.array_equal(R.x, np.minimum(d, t)))
    if args.cpu_scale > 0:
        line["cpu_baseline"] = cpu_reference(operands, min(args.cpu_scale, args.scale), line["cases"])
    out = json.dumps(line)
    print(out)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        open(args.out, "w").write(out + "\n")


if __name__ == "__main__":
    main()
