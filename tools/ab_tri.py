"""tools/ab_tri.py -- A/B of environment-selected variants of the masked dot (triangle counting,
C<L>=L*U' PLUS_TIMES_INT64 on RMAT) in ONE process on one GPU: every variant must return the same T
bit for bit as the first one (the baseline, itself parity-tested against the reference by
tests/), and its device / semiring-kernel times are printed.  Not a bench: no JSON contract.

    python tools/ab_tri.py --scale 22 --out gpurun_out/ab/ab.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

VARIANTS = [
    ("default", {}),
    ("old", {"GB200_DOTR": "0"}),
    ("notiny", {"GB200_DOTR_TINY": "0"}),
    ("valued_old", {"GB200_DOTR": "0", "GB200_DOTG_ISO": "0"}),
    ("bm_half", {"GB200_DOTR_BM_BITS": "753664"}),
    ("notrim", {"GB200_DOTG_TRIM": "0"}),
    ("chunk256", {"GB200_DOTG_CHUNK": "256"}),
    ("chunk512", {"GB200_DOTG_CHUNK": "512"}),
    ("chunk2048", {"GB200_DOTG_CHUNK": "2048"}),
    ("hub2048", {"GB200_DOTG_HUB_CHUNK": "2048"}),
    ("hub4096", {"GB200_DOTG_HUB_CHUNK": "4096"}),
    ("hub8192", {"GB200_DOTG_HUB_CHUNK": "8192"}),
    ("hub16384", {"GB200_DOTG_HUB_CHUNK": "16384"}),
    ("valued", {"GB200_DOTG_ISO": "0"}),
    ("valued_notrim", {"GB200_DOTG_ISO": "0", "GB200_DOTG_TRIM": "0"}),
    # round 2, second half: side streams, short-list trim skip
    ("nostreams", {"GB200_DOT_STREAMS": "0"}),
    ("trim1", {"GB200_DOTG_TRIM": "1"}),
    ("nostreams_trim1", {"GB200_DOT_STREAMS": "0", "GB200_DOTG_TRIM": "1"}),
    # the same T by the reference's other triangle-counting formulation: C<L> = L*L, masked saxpy
    # (the "outer product" lines of BASELINE.md 1); SAXPY is read by this script, not by the library
    ("masked_saxpy", {"SAXPY": "1"}),
]
KEYS = ("GB200_DOTG_TRIM", "GB200_DOTG_HUB_CHUNK", "GB200_DOTG_CHUNK", "GB200_DOTG_ISO", "GB200_DOTR",
        "GB200_DOTR_BM_BITS", "GB200_DOTR_TINY", "GB200_DOT_STREAMS")
# the answer of the first measured run (profiles/r1_trim): a base that is itself wrong is noticed
KNOWN = {22: (44374678, 2111700731)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--ef", type=int, default=16)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="", help="comma-separated variant names")
    args = ap.parse_args()
    t0 = time.time()
    import torch
    import bench
    import graphblas_b200 as gb
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    g = bench.build_rmat(args.scale, args.ef, dev)
    (n, Lp, Li, Lx), (_, Up, Ui, Ux) = bench.tri_operands(g)
    del g
    if dev == "cuda":
        torch.cuda.empty_cache()
    L = gb.Matrix(n, n, Lp, Li, Lx, None, "INT64")
    U = gb.Matrix(n, n, Up, Ui, Ux, None, "INT64")
    sr = gb.Semiring("PLUS", "TIMES", "INT64", flipxy=True)
    dL, dU = gb.DMatrix(L), gb.DMatrix(U)
    print(f"inputs ready after {time.time() - t0:.1f} s: n={n} nnz(L)={L.nnz}", flush=True)
    report = {"scale": args.scale, "n": n, "nnz_L": L.nnz, "variants": []}
    ref = None
    only = [v for v in args.only.split(",") if v]
    for name, env in VARIANTS:
        if only and name not in only:
            continue
        for k in KEYS:
            os.environ.pop(k, None)
        saxpy = env.get("SAXPY") == "1"
        os.environ.update({k: v for k, v in env.items() if k != "SAXPY"})
        row = {"name": name, "env": env}
        try:
            ms, kms = [], []
            mult = (lambda fetch: gb.axb_device(dL, False, dL, dL, sr, False, fetch=fetch)) if saxpy else \
                   (lambda fetch: gb.axb_device(dL, False, dU, dL, sr, True, fetch=fetch))
            for _ in range(args.reps):
                info = mult(False).info
                ms.append(info["device_ms"])
                kms.append(info["kernel_ms"])
            res = mult(True)
            T = res.matrix
            row.update(device_ms=ms, kernel_ms=kms, best_ms=min(ms), nnz_T=T.nnz,
                       madds=res.info["flops"], ntri=int(T.x.sum()))
            if args.scale in KNOWN and args.ef == 16:
                row["known_answer"] = bool((T.nnz, int(T.x.sum())) == KNOWN[args.scale])
            if ref is None:
                ref = T
                row["same_as_base"] = True
            else:
                row["same_as_base"] = bool(np.array_equal(ref.p, T.p) and np.array_equal(ref.i, T.i)
                                           and np.array_equal(ref.x, T.x))
        except Exception as e:      # a variant that fails must not hide the others
            row["error"] = repr(e)
        report["variants"].append(row)
        print(json.dumps(row), flush=True)
        if args.out:
            os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
            with open(args.out, "w") as f:
                json.dump(report, f, indent=1)
    print(f"total {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
