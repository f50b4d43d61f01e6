"""tools/transpose_bench.py -- C = A' on one GPU at the headline scale (the step
GB_AxB_meta runs in front of the multiply for a transposed operand, reference Source/GB_transpose.c).

L = tril (A,-1) of the RMAT graph of bench.py (scale 22, edge factor 16: 64 M entries, INT64 values), resident
in HBM; timed: gb200_transpose_device (CUDA events of the library, best and median of --reps after 2
warm-ups).  Checked, bit for bit and outside the timed region: L' == U = triu (A,1) of the symmetric graph
(pattern), (L')' == L (pattern and values), and at --check-scale the whole T against the oracle.
Roofline: algorithmic bytes = both matrices touched once in the API layout (8-byte pointers and indices),
2 * (8 * (n+1) + nnz * (8 + 8)); what the radix sort really moves is in the line (`moved_bytes`: three passes
of 20 bytes per entry, the position -> vector table, the gather).  Prints one JSON line; not the driver's bench
contract (bench.py), a measurement of a neighbour of the path.

    python tools/transpose_bench.py --scale 22 --check-scale 14 --out gpurun_out/tr/transpose_s22.json
"""
import argparse
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cpu_reference(bench, scale, ef, dev, gpu_ms, gpu_nnz):
    """the compiled reference's own GB_transpose (oracle/_ref, one thread: it is sequential,
    Source/GB_transpose_bucket.c) on L at `scale`, timed alone"""
    import ctypes as C
    import time
    import grbref
    if not grbref.available():
        return {"unavailable": "oracle/_ref not built"}
    G = grbref.GraphBLAS.get(with_shim=False)
    g = bench.build_rmat(scale, ef, dev)
    (n, Lp, Li, Lx), _ = bench.tri_operands(g)
    a = G.matrix_import("CSC", "INT64", n, n, Lp, Li, Lx)
    fn = G.lib.GB_transpose
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_bool, C.c_void_p, C.c_void_p, C.c_void_p]
    best = None
    for _ in range(2):
        T = C.c_void_p()
        t0 = time.perf_counter()
        G.ok(fn(C.byref(T), None, True, a, None, None), "GB_transpose")
        dt = time.perf_counter() - t0
        G.matrix_free(T)
        best = dt if best is None else min(best, dt)
    G.matrix_free(a)
    nnz = int(Lp[-1])
    return {"value": best * 1e3, "unit": "ms", "cores": 1, "kind": "reference",
            "sample": f"GB_transpose of L at scale {scale} ({nnz} entries), best of 2",
            "ns_per_entry": best * 1e9 / nnz, "gpu_ns_per_entry": gpu_ms * 1e6 / gpu_nnz}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--ef", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check-scale", type=int, default=14)
    ap.add_argument("--out", default="")
    ap.add_argument("--cpu-scale", type=int, default=20, help="the reference's own GB_transpose on the host, on "
                    "the same L at this scale (0: skip)")
    ap.add_argument("--ab", action="store_true", help="also time the variants the environment switches select "
                    "(GB200_TR_STAGED=0: direct scatter; GB200_TR_ISO=0: values gathered even for a pattern)")
    args = ap.parse_args()
    import graphblas_b200 as gb
    import bench
    import oracle_c
    import torch
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"      # cpu: only under tools/emu_*.py
    gb.init(0)
    line = {"metric": "GB_transpose on the device", "scale": args.scale, "edgefactor": args.ef}

    # parity at a size the oracle runs at: the whole T, form included
    g = bench.build_rmat(args.check_scale, args.ef, dev, weighted=True)
    n = g["n"]
    A = gb.Matrix(n, n, g["p"].cpu().numpy(), g["cols"].cpu().numpy(), g["w"].cpu().numpy(), None, "FP64")
    for ctype in (None, "FP32", "INT32"):
        ref = oracle_c.transpose(A, ctype)
        got = gb.transpose_host(A, ctype, hyper=False, hyper_ratio=0.0625).matrix
        same = ((ref.h is None) == (got.h is None) and np.array_equal(ref.p, got.p)
                and np.array_equal(ref.i, got.i) and np.array_equal(ref.x, got.x) and ref.type == got.type)
        if not same:
            raise SystemExit(f"transpose differs from the oracle at scale {args.check_scale}, ctype {ctype}")
    line["parity"] = {"oracle_scale": args.check_scale, "identical": True}

    g = bench.build_rmat(args.scale, args.ef, dev)
    (n, Lp, Li, Lx), (_, Up, Ui, Ux) = bench.tri_operands(g)
    del g
    L = gb.Matrix(n, n, Lp, Li, Lx, None, "INT64")
    dL = gb.DMatrix(L)
    ms = []
    for r in range(args.reps + 2):
        res = gb.transpose_device(dL, None, hyper=False, fetch=False)
        if r >= 2:
            ms.append(res.info["device_ms"])
    T = gb.transpose_device(dL, None, hyper=False, fetch=True).matrix
    if not (np.array_equal(T.p, Up) and np.array_equal(T.i, Ui)):
        raise SystemExit("tril (A)' != triu (A) of the symmetric graph")
    if args.ab:
        variants = {}
        for name, env in (("direct_scatter", {"GB200_TR_STAGED": "0"}), ("no_iso_shortcut", {"GB200_TR_ISO": "0"}),
                          ("round_start", {"GB200_TR_STAGED": "0", "GB200_TR_ISO": "0"})):
            os.environ.update(env)
            t = []
            for r in range(args.reps + 1):
                res = gb.transpose_device(dL, None, hyper=False, fetch=(r == 0))
                if r == 0:
                    same = np.array_equal(res.matrix.p, T.p) and np.array_equal(res.matrix.i, T.i) \
                        and np.array_equal(res.matrix.x, T.x)
                else:
                    t.append(res.info["device_ms"])
            for k in env:
                del os.environ[k]
            variants[name] = {"env": env, "best_ms": min(t), "identical_T": bool(same)}
        line["variants"] = variants
    dT = gb.DMatrix(T)
    TT = gb.transpose_device(dT, None, hyper=False, fetch=True).matrix
    if not (np.array_equal(TT.p, Lp) and np.array_equal(TT.i, Li) and np.array_equal(TT.x, Lx)):
        raise SystemExit("(L')' != L")
    nnz = int(Lp[-1])
    bits = max(1, int(np.ceil(np.log2(n))))
    passes = (bits + 7) // 8
    algo = 2 * (8 * (n + 1) + nnz * 16)
    moved = nnz * (passes * 20 + 4 + 4 + 8 + 4 + 8 + 4 + 8 + 1 + 1 + 8)
    peak_gbs, peak_source = bench.measured_peak()
    best = min(ms)
    line.update(n=n, nnz=nnz, radix_passes=passes, device_ms=ms, best_ms=best, median_ms=float(np.median(ms)),
                properties={"tril_transposed_is_triu": True, "twice_is_identity": True},
                roofline={"bound": "hbm", "achieved": algo / best / 1e6, "peak": peak_gbs, "unit": "GB/s",
                          "frac": algo / best / 1e6 / peak_gbs, "algorithmic_bytes": algo,
                          "moved_bytes_estimate": moved, "peak_source": peak_source})
    if args.cpu_scale > 0:
        line["cpu_baseline"] = cpu_reference(bench, min(args.cpu_scale, args.scale), args.ef, dev, best, nnz)
    out = json.dumps(line)
    print(out)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        open(args.out, "w").write(out + "\n")


if __name__ == "__main__":
    main()
