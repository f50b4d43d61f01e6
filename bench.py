#!/usr/bin/env python
"""bench.py -- the headline measurement: masked GrB_mxm on an RMAT graph, semiring GFLOP/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload tri|spgemm|spgemm_rmat|sssp|bfs]
  python bench.py --impl reference ...      # the reference's own CPU GrB_mxm, bounded sample

A "step" is one pass of the hot path (one GrB_mxm-equivalent multiply through the C ABI of
libgb_b200.so) over one synthetic input.  Default workload (BASELINE.json: "GrB_mxm ... RMAT
scale-22"): `tri` = C<L> = L*U' over PLUS_TIMES_INT64 on RMAT scale 22, edge factor 16 -- what
GrB_mxm(C, L, NULL, GxB_PLUS_TIMES_INT64, L, U, desc{INP1=TRAN}) hands to GB_AxB_parallel
(reference Demo/Source/tricount.c:166-178): M = L, A = U, B = L, do_adotb, flipxy.

  parity : BEFORE any number is printed the GPU's T is compared with the compiled reference
          (oracle/_ref) on the vectors the cpu_baseline sample computes: identical pointers, pattern
          and values (integer / bool / MIN / MAX: bit-exact; fp64 PLUS_TIMES: 64 eps in the 1-norm,
          the reference's own test criterion).  A mismatch exits non-zero with no JSON line.
  value : 2 * madds / t, operands resident in HBM (gb200_AxB_device); t = CUDA-event time of the K
          steps on the library's launching stream (gb200_timer_mark), max over ranks
  e2e   : same metric through gb200_AxB_host + gb200_result_fetch with HOST buffers in page-locked
          memory from gb200_host_malloc (what every GraphBLAS array is after GxB_init with the
          gb200_host_* allocator): H2D of M, A, B and D2H of T inside the timed region
  t_api : wall time of the UNMODIFIED reference's GrB_mxm / GrB_mxv / GrB_vxm with
          libgb_b200_shim.so interposed (GB_AxB_parallel -> GPU), GraphBLAS started with
          GxB_init (gb200_host_*): the call an application makes (SURVEY.md 8d "Timing window")
  roofline : dominant kernel (the semiring kernel: dot / saxpy), algorithmic bytes of SURVEY.md
          8(d) / its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs; `traffic` = DRAM bytes of
          the same kernels from the committed ncu capture (profiles/traffic.json)
  cpu_baseline : oracle/_ref (the compiled reference), GrB_mxm / GrB_mxv on a bounded sample of
          the same workload, on rank 0 at N == 1 only.  One call of the reference is sequential
          (Source/GB_AxB_parallel.c:102-103), so the sample is cut into one independent slice of
          output vectors per host core and the slices run concurrently from user threads (the
          reference is thread-safe for that, Demo/Program/pthread_demo.c)
  secondary : the same {value, roofline, e2e, t_api, parity, cpu_baseline} for the other configs of
          BASELINE.json on one GPU: unmasked C=A*A (saxpy) on RMAT, SSSP (mxv) and BFS (vxm) at scale 22
  neighbours : lines of the components next to the path, each measured by its own tool in a
          process of its own after everything else: the device transpose (row f2, tools/transpose_bench.py)
          and the device accum / mask step (row f1, tools/accum_mask_bench.py)

N > 1 (torchrun, one rank per GPU): the mask's entries are split into N owner-aligned parts (the
reference's own plan, GB_AxB_parallel.c:52); A and B are replicated; no data-path collective; the
triangle count is all-reduced over NCCL.  The problem is fixed, so scaling is "strong".

`--impl reference` never imports graphblas_b200 (so no product library is mapped into that process)
and builds its inputs on the CPU with the same counter-based generator.
"""
from __future__ import annotations

import argparse
import copy
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "GrB_mxm semiring GFLOP/s"


def pure_containers():
    """graphblas_b200/containers.py loaded BY PATH: Matrix / Semiring without the package's
    __init__ (which maps libgb_b200.so).  Used by the reference arm."""
    name = "gb200_pure_containers"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(ROOT, "graphblas_b200", "containers.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


# ---------------------------------------------------------------------------------------------
# inputs
# ---------------------------------------------------------------------------------------------
def build_rmat(scale: int, ef: int, device: str, weighted: bool = False):
    """-> dict of torch tensors: n, rows, cols (symmetric, loop-free, sorted), p, optional w"""
    import gen
    n, r, c = gen.rmat_edges(scale, ef, 42, device)
    p = gen.csr_from_sorted(n, r, c)
    out = {"n": n, "rows": r, "cols": c, "p": p}
    if weighted:
        out["w"] = gen.rmat_weights(r, c)
    return out


def tri_operands(g, dtype=np.int64):
    """L = tril(A,-1), U = triu(A,1) as host CSR arrays (vectors are rows)"""
    import gen
    n, r, c = g["n"], g["rows"], g["cols"]
    low = c < r
    Lr, Lc = r[low], c[low]
    Ur, Uc = r[~low], c[~low]
    Lp = gen.csr_from_sorted(n, Lr, Lc).cpu().numpy()
    Up = gen.csr_from_sorted(n, Ur, Uc).cpu().numpy()
    Li, Ui = Lc.cpu().numpy(), Uc.cpu().numpy()
    return (n, Lp, Li, np.ones(len(Li), dtype=dtype)), (n, Up, Ui, np.ones(len(Ui), dtype=dtype))


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in-process every 10 ms
    (nvidia_ml_py), else one nvidia-smi query per 100 ms."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.samples, self.stop, self.t = index, [], False, None
        self.nvml = self.handle = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)]
        return [str(mhz), str(self.max_mhz)] + ["Active" if (r & b) else "Not Active" for b in bits]

    def _run(self):
        while not self.stop:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    time.sleep(0.01)
                    continue
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for k, nm in enumerate(self.NAMES):
                if len(s) > 2 + k and s[2 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": sorted(reasons), "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# workloads: each returns the operands at the GB_AxB_parallel seam
# ---------------------------------------------------------------------------------------------
def algo_bytes(mats, cnvec, cnz, zsize):
    """API layout, 8-byte indices, every operand touched once"""
    b = 8 * (cnvec + 1) + cnz * (8 + zsize)
    for m in mats:
        b += 8 * (len(m.p)) + m.nnz * (8 + m.x.dtype.itemsize)
        if m.h is not None:
            b += 8 * len(m.h)
    return b


def make_workload(C, args, device):
    """C: the containers module (graphblas_b200 itself, or the pure containers for the reference arm)"""
    w = {}
    if args.workload == "tri":
        g = build_rmat(args.scale, args.ef, device)
        (n, Lp, Li, Lx), (_, Up, Ui, Ux) = tri_operands(g)
        L = C.Matrix(n, n, Lp, Li, Lx, None, "INT64")
        U = C.Matrix(n, n, Up, Ui, Ux, None, "INT64")
        w.update(name=f"GrB_mxm C<L>=L*U' PLUS_TIMES_INT64 (masked dot), RMAT scale {args.scale} "
                      f"edgefactor {args.ef}, n={n}, nnz(L)={L.nnz}",
                 M=L, A=U, B=L, mask_comp=False, do_adotb=True,
                 semiring=C.Semiring("PLUS", "TIMES", "INT64", flipxy=True), dtype="int64",
                 slice="M", kernel="dotr_kernel x4 + dotr_warp_kernel x2 + dot_kernel (masked dot)")
    elif args.workload == "spgemm":
        import gen
        n = 1 << args.scale
        A = gen.er(n, n, args.ef * n, 1)
        B = gen.er(n, n, args.ef * n, 2)
        Am, Bm = C.Matrix.from_scipy(A, "FP64"), C.Matrix.from_scipy(B, "FP64")
        # CSR C=A*B: the seam sees A := B_in, B := A_in, flipxy
        w.update(name=f"GrB_mxm C=A*B PLUS_TIMES_FP64 (saxpy), Erdos-Renyi n=2^{args.scale}, "
                      f"{args.ef} nnz/row",
                 M=None, A=Bm, B=Am, mask_comp=False, do_adotb=False,
                 semiring=C.Semiring("PLUS", "TIMES", "FP64", flipxy=True), dtype="f64", slice="B")
    elif args.workload == "spgemm_rmat":
        g = build_rmat(args.scale, args.ef, device, weighted=True)
        n = g["n"]
        Am = C.Matrix(n, n, g["p"].cpu().numpy(), g["cols"].cpu().numpy(), g["w"].cpu().numpy(), None,
                      "FP64")
        w.update(name=f"GrB_mxm C=A*A PLUS_TIMES_FP64 (saxpy), RMAT scale {args.scale} edgefactor "
                      f"{args.ef}, n={n}, nnz(A)={Am.nnz}",
                 M=None, A=Am, B=Am, mask_comp=False, do_adotb=False,
                 semiring=C.Semiring("PLUS", "TIMES", "FP64", flipxy=True), dtype="f64", slice="B")
    elif args.workload == "sssp":
        # GrB_mxv (d, NULL, GrB_MIN_FP64, GxB_MIN_PLUS_FP64, A, d, NULL), A CSR: the seam sees
        # A'*B by dot products with B = d (n x 1, all entries present), SURVEY.md 3.5
        g = build_rmat(args.scale, args.ef, device, weighted=True)
        n = g["n"]
        Am = C.Matrix(n, n, g["p"].cpu().numpy(), g["cols"].cpu().numpy(), g["w"].cpu().numpy(), None,
                      "FP64")
        src = int(np.argmax(np.diff(Am.p)))
        d = np.full(n, np.inf)
        d[src] = 0.0
        dv = C.Matrix(n, 1, np.array([0, n]), np.arange(n), d, None, "FP64")
        w.update(name=f"GrB_mxv d=A min.+ d MIN_PLUS_FP64 (dot, dense vector), RMAT scale {args.scale} "
                      f"edgefactor {args.ef}, n={n}, nnz(A)={Am.nnz}, one Bellman-Ford relaxation",
                 M=None, A=Am, B=dv, mask_comp=False, do_adotb=True,
                 semiring=C.Semiring("MIN", "PLUS", "FP64"), dtype="f64", slice="A",
                 kernel="spmv_stream_kernel")
    elif args.workload == "bfs":
        # bfs5m level loop (Demo/Source/bfs5m.c:71-82): q<!v> = q*A over LOR_LAND_BOOL, A CSR: the
        # seam sees A*B by saxpy with B = q (n x 1), M = v complemented, flipxy
        g = build_rmat(args.scale, args.ef, device)
        n = g["n"]
        Am = C.Matrix(n, n, g["p"].cpu().numpy(), g["cols"].cpu().numpy(),
                      np.ones(len(g["cols"]), dtype=np.bool_), None, "BOOL")
        w.update(name=f"BFS level loop GrB_vxm q<!v>=q*A LOR_LAND_BOOL (saxpy, vector), RMAT scale "
                      f"{args.scale} edgefactor {args.ef}, n={n}, nnz(A)={Am.nnz}",
                 M=None, A=Am, B=None, mask_comp=True, do_adotb=(args.bfs_dir == "pull"),
                 semiring=C.Semiring("LOR", "LAND", "BOOL", flipxy=True), dtype="bool", slice="none",
                 bfs_source=int(np.argmax(np.diff(Am.p))),
                 kernel="dotv_kernel (pull)" if args.bfs_dir == "pull" else "saxpyv_kernel + saxpyv_long_kernel (push)")
    else:
        raise SystemExit(f"unknown workload {args.workload}")
    sr = w["semiring"]
    w["config"] = {"workload": w["name"], "scale": args.scale, "edgefactor": args.ef,
                   "semiring": f"{sr.add}_{sr.mult}_{sr.xytype}",
                   "l2": "inputs larger than L2 (no flush needed)"
                         if sum(x.nnz for x in (w["A"],) if x is not None) * 12 > 2.6e8
                         else "inputs smaller than L2"}
    return w


def bfs_levels(gb, w, dA):
    """run the level loop once (through the device entry point) and record, per level, the frontier
    q and the visited vector v that the reference would hand to GB_AxB_parallel"""
    A = w["A"]
    n = A.vlen
    visited = np.zeros(n, dtype=bool)
    q = np.array([w["bfs_source"]], dtype=np.int64)
    levels = []
    while len(q):
        visited[q] = True
        vi = np.nonzero(visited)[0]
        qm = gb.Matrix(n, 1, np.array([0, len(q)]), q, np.ones(len(q), np.bool_), None, "BOOL")
        vm = gb.Matrix(n, 1, np.array([0, len(vi)]), vi, np.ones(len(vi), np.bool_), None, "BOOL")
        levels.append((qm, vm))
        r = gb.axb_device(gb.DMatrix(vm), True, dA, gb.DMatrix(qm), w["semiring"], w["do_adotb"])
        t = r.matrix.i
        q = t[~visited[t]] if not r.info["mask_applied"] else t
    return levels


def slice_vectors(C, m, lo, hi):
    """keep vectors [lo,hi) of a standard-form matrix, same dimensions (others become empty)"""
    p = np.zeros(len(m.p), dtype=np.int64)
    s, e = m.p[lo], m.p[hi]
    p[lo:hi + 1] = m.p[lo:hi + 1] - s
    p[hi + 1:] = e - s
    return C.Matrix(m.vlen, m.vdim, p, m.i[s:e], m.x[s:e], None, m.type)


# ---------------------------------------------------------------------------------------------
# the reference's public API on the workload: cpu_baseline / reference arm (shim off) and
# t_api (shim on)
# ---------------------------------------------------------------------------------------------
def _interleaved_parts(m, stride, nparts):
    """Every `stride`-th vector of m, dealt round-robin to `nparts` slices.  Each slice keeps the
    dimensions of m (the other vectors are empty).  -> list of (p, i, x), list of kept-vector masks"""
    nvec = m.nvec
    if stride == 1 and nparts == 1:
        return [(m.p, m.i, m.x)], [np.ones(nvec, dtype=bool)]
    cnt = np.diff(m.p)
    v = np.arange(nvec)
    sampled = (v % stride) == 0
    owner = (v // stride) % nparts
    parts, keeps = [], []
    for t in range(nparts):
        keep = sampled & (owner == t)
        cnt2 = np.where(keep, cnt, 0)
        p2 = np.concatenate([[0], np.cumsum(cnt2)])
        sel = np.repeat(keep, cnt)
        parts.append((p2, m.i[sel], m.x[sel]))
        keeps.append(keep)
    return parts, keeps


def host_bfs_levels(A, src):
    """(q, visited) index arrays per level of the bfs5m loop, computed with numpy on the host (used
    by the reference arm, which must not touch the GPU library)"""
    n = A.vlen
    visited = np.zeros(n, dtype=bool)
    q = np.array([src], dtype=np.int64)
    out = []
    while len(q):
        visited[q] = True
        out.append((q, np.nonzero(visited)[0]))
        starts, ends = A.p[q], A.p[q + 1]
        tot = int((ends - starts).sum())
        if tot == 0:
            break
        idx = np.repeat(starts - np.concatenate([[0], np.cumsum(ends - starts)[:-1]]), ends - starts) \
            + np.arange(tot)
        nb = np.unique(A.i[idx])
        q = nb[~visited[nb]]
    return out


def api_run(G, wl, w, stride, T, keep=False, reps=1):
    """The workload through the reference library's PUBLIC API (GrB_mxm / GrB_mxv / GrB_vxm) on every
    `stride`-th output vector, cut into T slices that run concurrently on T user threads.  With the
    shim off this is the reference's CPU path (cpu_baseline, reference arm); with the shim on and
    stride = T = 1 it is the call an application makes, landing on the GPU (t_api).
    -> dict(seconds (best of reps), madds, desc, threads, outputs (keep=True: what the calls
    returned, exported, with the kept-vector masks))"""
    from concurrent.futures import ThreadPoolExecutor
    import grbref
    sr = w["semiring"]
    sr_name = "GxB_" + "_".join((sr.add, sr.mult, sr.xytype))
    tname = sr.xytype
    handles, vhandles = [], []
    outputs = None
    if wl == "tri":
        # user-level call: C<Ls> = L*U', CSR, desc INP1 = TRAN (Demo/Source/tricount.c:166-178)
        L, U = w["B"], w["A"]
        n = L.vdim
        parts, keeps = _interleaved_parts(L, stride, T)
        l = G.matrix_import("CSR", tname, n, n, L.p, L.i, L.x)
        u = G.matrix_import("CSR", tname, n, n, U.p, U.i, U.x)
        ms = [l] if (stride == 1 and T == 1) else [G.matrix_import("CSR", tname, n, n, *pt) for pt in parts]
        d = G.descriptor(inp1=grbref.GrB_TRAN)
        handles = [l, u] + (ms if ms[0] is not l else [])
        cs = []

        def prepare():
            cs[:] = [G.matrix_new(tname, n, n) for _ in parts]

        ntri = [0] * len(parts)

        def run(t):
            # the triangle count as the reference's tricount does it (Demo/Source/tricount.c:166-178):
            # the masked multiply, then GrB_reduce of C over PLUS_INT64
            G.mxm(cs[t], ms[t], None, sr_name, l, u, d)
            ntri[t] = G.reduce_int64(cs[t])
        madds_of = lambda: sum(ntri)                # every match is one multiply-add
        what = f"every {stride}th vector of the mask L ({int(sum(k.sum() for k in keeps))} of {L.nvec})"
        call = "GrB_mxm + GrB_reduce"
        export = lambda: [G.matrix_export(c, "CSR") for c in cs]
        release = lambda: [G.matrix_free(c) for c in cs]
    elif wl in ("spgemm", "spgemm_rmat"):
        # user-level call: C = As*B, CSR
        Ain, Bin = w["B"], w["A"]
        parts, keeps = _interleaved_parts(Ain, stride, T)
        b = G.matrix_import("CSR", tname, Bin.vdim, Bin.vlen, Bin.p, Bin.i, Bin.x)
        As = [G.matrix_import("CSR", tname, Ain.vdim, Ain.vlen, *pt) for pt in parts]
        handles = [b] + As
        lenB = np.diff(Bin.p)
        cs = []

        def prepare():
            cs[:] = [G.matrix_new(tname, Ain.vdim, Bin.vlen) for _ in parts]

        def run(t):
            G.mxm(cs[t], None, None, sr_name, As[t], b, None)
            G.matrix_nvals(cs[t])
        madds_of = lambda: int(sum(int(lenB[pt[1]].sum()) for pt in parts))
        what = f"every {stride}th row of A ({int(sum(k.sum() for k in keeps))} of {Ain.nvec})"
        call = "GrB_mxm"
        export = lambda: [G.matrix_export(c, "CSR") for c in cs]
        release = lambda: [G.matrix_free(c) for c in cs]
    elif wl == "sssp":
        # user-level call: w = As min.+ d (GrB_mxv, A CSR): rows of A are independent outputs
        Am, dv = w["A"], w["B"]
        n = Am.vdim
        parts, keeps = _interleaved_parts(Am, stride, T)
        As = [G.matrix_import("CSR", tname, n, n, *pt) for pt in parts]
        dd = G.vector_import(tname, n, dv.i, dv.x)
        handles, vhandles = As, [dd]
        ws = []

        def prepare():
            ws[:] = [G.vector_new(tname, n) for _ in parts]

        def run(t):
            G.mxv(ws[t], None, None, sr_name, As[t], dd, None)
            G.vector_nvals(ws[t])
        madds_of = lambda: int(sum(len(pt[1]) for pt in parts))     # d is dense: every entry matches
        what = f"every {stride}th row of A ({int(sum(k.sum() for k in keeps))} of {Am.nvec})"
        call = "GrB_mxv"
        export = lambda: [G.vector_export(x) for x in ws]
        release = lambda: [G.vector_free(x) for x in ws]
    else:
        # bfs: the level loop q<!v> = q*A (GrB_vxm, REPLACE, SCMP); one frontier vector cannot be
        # sliced by the caller, so this leg is one thread
        T = 1
        Am = w["A"]
        n = Am.vdim
        levels = w.get("host_levels") or host_bfs_levels(Am, w["bfs_source"])
        w["host_levels"] = levels
        keeps = None
        a = G.matrix_import("CSR", tname, n, n, Am.p, Am.i, Am.x)
        qs = [G.vector_import(tname, n, q, np.ones(len(q), np.bool_)) for q, _ in levels]
        vs = [G.vector_import(tname, n, v, np.ones(len(v), np.bool_)) for _, v in levels]
        d = G.descriptor(outp=grbref.GrB_REPLACE, mask=grbref.GrB_SCMP)
        handles, vhandles = [a], qs + vs
        lenA = np.diff(Am.p)
        ws = []

        def prepare():
            ws[:] = [G.vector_new(tname, n) for _ in levels]

        def run(t):
            for q, v, wv in zip(qs, vs, ws):
                G.vxm(wv, v, None, sr_name, q, a, d)
                G.vector_nvals(wv)
        madds_of = lambda: int(sum(int(lenA[q].sum()) for q, _ in levels))
        what = f"all {len(levels)} levels"
        call = "GrB_vxm"
        export = lambda: [G.vector_export(x) for x in ws]
        release = lambda: [G.vector_free(x) for x in ws]

    best = None
    for rep in range(reps):
        prepare()
        t0 = time.perf_counter()
        if T == 1:
            run(0)
        else:
            with ThreadPoolExecutor(T) as ex:
                list(ex.map(run, range(T)))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        if rep + 1 < reps:
            release()
    madds = madds_of()
    if keep:
        outputs = {"parts": export(), "keeps": keeps}      # export is destructive: frees the outputs
    else:
        release()
    for h in handles:
        G.matrix_free(h)
    for h in vhandles:
        G.vector_free(h)
    desc = (f"{what}, cut into {T} slices of output vectors, one {call} per slice on {T} concurrent "
            f"user threads of {os.cpu_count()} host cores (a single reference call is sequential: "
            f"Source/GB_AxB_parallel.c:102-103)")
    return {"seconds": best, "madds": madds, "gflops": 2.0 * madds / best / 1e9, "desc": desc,
            "threads": T, "outputs": outputs}


def pick_stride(wl, madds_total, T, seconds=12.0):
    """sample every n-th output vector so that the reference's CPU leg takes about `seconds`:
    measured rates of the reference on one host thread, madds/s (profiles/r1)"""
    rate = {"tri": 6.0e6, "spgemm": 2.0e7, "spgemm_rmat": 3.0e7, "sssp": 8.0e7, "bfs": 1e12}[wl]
    return int(max(1, np.ceil(madds_total / (rate * T * seconds))))


def _rows_of(p, i, x, keep):
    """entries of the kept vectors of a CSR-like triple, in order -> (counts, i, x)"""
    cnt = np.diff(p)
    sel = np.repeat(keep, cnt)
    return cnt[keep], i[sel], x[sel]


def parity_check(wl, w, sample, gpu_T):
    """The vectors the reference computed (sample['outputs']) against the same vectors of the GPU's T.
    gpu_T: a Matrix (tri / spgemm / sssp) or the list of per-level index arrays (bfs).
    Integer / bool / MIN-PLUS: identical.  fp64 PLUS_TIMES: identical pattern, values within 64 eps in
    the 1-norm (Test/GB_spec_compare.m:18-24).  -> the `parity` object of the JSON line"""
    out = {"against": "oracle/_ref (the compiled reference) through GrB_mxm/mxv/vxm on the "
                      "cpu_baseline sample", "checked_vectors": 0, "checked_entries": 0,
           "identical": True, "criterion": "bit-exact pointers, pattern, values"}
    fp_plus = (w["semiring"].add in ("PLUS", "TIMES") and w["semiring"].xytype in ("FP32", "FP64"))
    if fp_plus:
        out["criterion"] = "identical pointers and pattern; values: |ref-got|_1 <= 64 eps |ref|_1"
    parts, keeps = sample["outputs"]["parts"], sample["outputs"]["keeps"]
    num = den = 0.0

    def fail(msg):
        out["identical"] = False
        out["first_difference"] = msg

    if wl in ("tri", "spgemm", "spgemm_rmat"):
        for ex, keep in zip(parts, keeps):
            rc, ri, rx = _rows_of(ex["Ap"], ex["Ai"], ex["Ax"], keep)
            gc, gi, gx = _rows_of(gpu_T.p, gpu_T.i, gpu_T.x, keep)
            out["checked_vectors"] += int(keep.sum())
            out["checked_entries"] += int(rc.sum())
            if int(np.diff(ex["Ap"])[~keep].sum()) != 0:
                return fail("the reference wrote outside the sampled vectors") or out
            if not np.array_equal(rc, gc):
                return fail("vector lengths differ") or out
            if not np.array_equal(ri, gi):
                return fail("patterns differ") or out
            if fp_plus:
                if not np.array_equal(np.isnan(rx), np.isnan(gx)) or \
                        not np.array_equal(rx[np.isinf(rx)], gx[np.isinf(rx)]):
                    return fail("NaN / Inf placement differs") or out
                fin = np.isfinite(rx)
                num += float(np.abs(rx[fin] - gx[fin]).sum())
                den += float(np.abs(rx[fin]).sum())
            elif not np.array_equal(rx, gx):
                return fail("values differ") or out
    elif wl == "sssp":
        gmap = np.full(gpu_T.vlen, -1, dtype=np.int64)
        gmap[gpu_T.i] = np.arange(gpu_T.nnz)
        for ex, keep in zip(parts, keeps):
            vi, vx = ex["vi"], ex["vx"]
            order = np.argsort(vi, kind="stable")
            vi, vx = vi[order], vx[order]
            want = np.nonzero(keep & (gmap >= 0))[0]
            out["checked_vectors"] += int(keep.sum())
            out["checked_entries"] += len(vi)
            if not np.array_equal(vi, want):
                return fail("pattern of w differs") or out
            if not np.array_equal(vx.view(np.uint64), gpu_T.x[gmap[vi]].view(np.uint64)):
                return fail("values of w differ") or out
    else:
        for lvl, (ex, got) in enumerate(zip(parts, gpu_T)):
            vi = np.sort(ex["vi"])
            out["checked_vectors"] += 1
            out["checked_entries"] += len(vi)
            if not np.array_equal(vi, np.sort(got)):
                return fail(f"frontier of level {lvl} differs") or out
            if not ex["vx"].all():
                return fail(f"level {lvl}: a false value in the reference's w") or out
        if len(parts) != len(gpu_T):
            return fail("number of levels differs") or out
    if fp_plus:
        out["rel_err_1norm"] = num / den if den > 0 else 0.0
        out["tolerance"] = 64 * float(np.finfo(np.float64).eps)
        if den > 0 and num > out["tolerance"] * den:
            fail("values outside 64 eps in the 1-norm")
    return out


# ---------------------------------------------------------------------------------------------
# our arm: one workload on this process' GPU (rank of world)
# ---------------------------------------------------------------------------------------------
def measure(args, env, primary=True):
    import torch
    import torch.distributed as dist
    import graphblas_b200 as gb
    rank, world, local_rank, device = env["rank"], env["world"], env["local_rank"], env["device"]
    w = make_workload(gb, args, device)
    torch.cuda.empty_cache()
    A, B, M = w["A"], w["B"], w["M"]
    sliced_name = w["slice"]
    dA = gb.DMatrix(A)
    do_cpu = (world == 1 and not args.no_cpu and rank == 0)

    # ---- the multiplies of one step: (M, A, B) host triples + their resident handles ------------
    slabs = None
    if args.workload == "bfs":
        levels = bfs_levels(gb, w, dA)
        nslices, srank = (args.slice_of, args.slice_rank) if (world == 1 and args.slice_of > 1) else (world, rank)
        if nslices > 1:
            # vector push on N GPUs: a rank owns a block of vertices = of A's vectors (balanced by their
            # entries) and pushes the frontier entries of its block; the partial frontiers meet in every
            # GPU's dense copy of w (gb200_peerbuf_publish: every rank stores the same value `true`, so
            # overlapping stores of LOR need no atomic), SURVEY.md 8e
            if w["do_adotb"]:
                raise SystemExit("bench.py shards the BFS push step; use --bfs-dir push for N > 1")
            bounds = gb.partition_by_flops(A.p, nslices)
            lo, hi = int(bounds[srank]), int(bounds[srank + 1])
            A_r = slice_vectors(gb, A, lo, hi)
            dA_r = gb.DMatrix(A_r)

            def q_slice(qm):
                qi = qm.i[(qm.i >= lo) & (qm.i < hi)]
                return gb.Matrix(A.vlen, 1, np.array([0, len(qi)]), qi, np.ones(len(qi), np.bool_), None, "BOOL")
            calls = []
            for qm, vm in levels:
                qr = q_slice(qm)
                calls.append((vm, A_r, qr, gb.DMatrix(vm), dA_r, gb.DMatrix(qr)))
            dA.free()
            w["partition"] = f"{nslices} blocks of A's vectors (vertices) balanced by entries; every level's frontier cut at the same bounds"
        else:
            calls = [(vm, A, qm, gb.DMatrix(vm), dA, gb.DMatrix(qm)) for qm, vm in levels]
        w["name"] += f", {len(levels)} levels from vertex {w['bfs_source']}"
        w["config"]["workload"] = w["name"]
    else:
        dB = dA if B is A else gb.DMatrix(B)
        dM = None
        if M is not None:
            dM = dB if M is B else gb.DMatrix(M)
        # flop-balanced contiguous slices of the sliced operand's vectors, one per rank
        nslices, srank = (args.slice_of, args.slice_rank) if (world == 1 and args.slice_of > 1) else (world, rank)
        if nslices > 1:
            sliced = {"M": M, "B": B, "A": A}[sliced_name]
            if sliced_name == "B":
                cum, _ = gb.flopcount(dM, dA, dB)
            elif sliced_name == "A":
                cum = A.p                       # dot with a vector: work of a vector of A = its length
            else:
                cum = None                      # masked dot: cut by owner vector, below
            if sliced_name == "M":
                # a rank takes the mask entries whose owner vector (the longer of A(:,i), B(:,j)) is
                # in its range: no hub is loaded by more than one rank (graphblas_b200/sharded.py).
                # The cut follows a cost model of the kernels; two calibration rounds (one multiply per
                # rank, outside the timed region) correct it with the ranks' measured times.
                from graphblas_b200 import sharded
                op = sharded.OwnerPartition(M, A, B, nslices)
                calib = []

                def part_ms(r):
                    dm_r = gb.DMatrix(op.mask(r))
                    t = 0.0
                    for _ in range(2):              # the first one warms the allocator
                        t = gb.axb_device(dm_r, w["mask_comp"], dA, dB, w["semiring"], w["do_adotb"],
                                          fetch=False).info["device_ms"]
                    dm_r.free()
                    return t
                for _ in range(args.calibrate):
                    if world > 1:
                        tl = torch.zeros(world, dtype=torch.float64, device=device)
                        tl[rank] = part_ms(rank)
                        dist.all_reduce(tl)
                        times = tl.tolist()
                    else:
                        times = [part_ms(r) for r in range(nslices)]   # --slice-of: every rank in turn
                    calib.append([round(v, 3) for v in times])
                    op.rebalance(times)
                w["calibration_ms"] = calib
                mine = op.mask(srank)
                bounds = None
            else:
                bounds = gb.partition_by_flops(cum, nslices)
                lo, hi = int(bounds[srank]), int(bounds[srank + 1])
                mine = slice_vectors(gb, sliced, lo, hi)
            dmine = gb.DMatrix(mine)
            if sliced_name == "M":
                M, dM = mine, dmine
            elif sliced_name == "A":
                A, dA = mine, dmine
            else:
                B, dB = mine, dmine
        calls = [(M, A, B, dM, dA, dB)]
        # ---- C too large for HBM: column slabs (SURVEY.md 7 "hard parts", BASELINE cfg 5) -------------
        # The flop count runs first (gb200_flopcount_device = GB_AxB_flopcount on the device): nnz (C) <=
        # flops, so when flops x 16 B (pattern + accumulators + symbolic workspace) exceeds the budget,
        # this rank's vectors of B are cut into flop-balanced slabs; every slab of C is computed, reduced
        # to a checksum on the device (gb200_result_reduce) and discarded.
        slabs = None
        if args.workload in ("spgemm", "spgemm_rmat") and M is None:
            cum_b, total_flops = gb.flopcount(None, dA, dB)
            free_b, _ = torch.cuda.mem_get_info()
            budget = int(args.slab_gb * 2 ** 30) if args.slab_gb > 0 else int(0.50 * free_b)
            nslab = int(-(-total_flops * 16 // budget))
            if nslab > 1:
                sb = gb.partition_by_flops(cum_b, nslab)
                parts = [slice_vectors(gb, B, int(sb[k]), int(sb[k + 1])) for k in range(nslab)
                         if sb[k + 1] > sb[k]]
                calls = [(None, A, bk, None, dA, gb.DMatrix(bk)) for bk in parts]
                if dB is not dA:
                    dB.free()
                slabs = {"count": len(calls), "flops_total": int(total_flops),
                         "budget_bytes": budget, "whole_B": B,
                         "rule": "flops x 16 B per slab <= budget (nnz (C) <= flops); every slab of C "
                                 "is reduced to a checksum on the device and discarded"}
                w["config"]["slabs"] = len(calls)

    # ---- parity gate (one GPU): the reference on a bounded sample, the GPU's T on the same vectors
    parity = cpu = None
    G = None
    if do_cpu or (world == 1 and not args.no_api and rank == 0):
        import grbref
        try:
            G = grbref.GraphBLAS.get(with_shim=True, pinned=True)
            G.use_gpu(False)
        except Exception as e:          # the reference .so is test infrastructure; say so if absent
            G = None
            cpu = {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "reference",
                   "sample": f"unavailable: {e}"}
    if do_cpu and G is not None:
        T = max(1, args.cpu_threads or (os.cpu_count() or 1))
        first = None
        if slabs is None:
            first = gb.axb_device(calls[0][3], w["mask_comp"], calls[0][4], calls[0][5], w["semiring"],
                                  w["do_adotb"], fetch=(args.workload != "bfs"))
        if slabs is not None:
            # the whole T never exists: the GPU computes exactly the output vectors the reference's
            # sample computes (one multiply with the other vectors of B emptied) and those are compared
            madds_total = slabs["flops_total"]
            stride = args.cpu_stride if args.cpu_stride > 0 else pick_stride(args.workload, madds_total, T)
            sample = api_run(G, args.workload, w, stride, T, keep=True)
            (sp_, si_, sx_), _ = _interleaved_parts(slabs["whole_B"], stride, 1)[0][0], None
            Bs = gb.Matrix(B.vlen, B.vdim, sp_, si_, sx_, None, B.type)
            gpu_T = gb.axb_host(None, False, A, Bs, w["semiring"], w["do_adotb"]).matrix
            del Bs
        elif args.workload == "bfs":
            madds_total = int(sum(int(np.diff(A.p)[qm.i].sum()) for qm, _ in levels))
            gpu_T = []
            for (vm, _, qm, dvm, da, dqm) in calls:
                r = gb.axb_device(dvm, True, da, dqm, w["semiring"], w["do_adotb"])
                t = r.matrix.i
                if not r.info["mask_applied"]:
                    vis = np.zeros(A.vlen, dtype=bool)
                    vis[vm.i] = True
                    t = t[~vis[t]]
                gpu_T.append(t)
            w["host_levels"] = [(qm.i, vm.i) for qm, vm in levels]
        else:
            madds_total = first.info["flops"]
            gpu_T = first.matrix
        if slabs is None:
            stride = args.cpu_stride if args.cpu_stride > 0 else pick_stride(args.workload, madds_total, T)
            sample = api_run(G, args.workload, w, stride, T, keep=True)
        parity = parity_check(args.workload, w, sample, gpu_T)
        cpu = {"value": sample["gflops"], "unit": "GFLOP/s", "cores": sample["threads"],
               "kind": "reference", "sample": sample["desc"], "seconds": sample["seconds"],
               "host_cores_present": os.cpu_count()}
        del sample, gpu_T, first
        if not parity["identical"]:
            sys.stderr.write(f"PARITY GATE FAILED ({args.workload}): {json.dumps(parity)}\n")
            sys.exit(3)

    # host operands of the end-to-end leg live in page-locked memory (gb200_host_malloc)
    pinned = {}

    def pin(m):
        if m is None:
            return None
        if id(m) not in pinned:
            pinned[id(m)] = m.pinned()
        return pinned[id(m)]
    if slabs is not None and world > 1:
        args.no_e2e = True          # the slab-streamed multiply on N GPUs is measured on resident operands
    if world > 1 and args.workload == "bfs":
        args.no_e2e = True          # the sharded level loop is measured on resident operands
    hcalls = [] if args.no_e2e else [(pin(m), pin(a), pin(b)) for (m, a, b, _, _, _) in calls]
    hfull = hcalls

    # vector pull on N > 1 GPUs: every rank owns a block of A's vectors = of w's entries, and the
    # slices of w are all-gathered over NCCL every step; T never leaves HBM
    exchange = None
    if world > 1 and ((sliced_name == "A" and A.vdim >= 1 and B is not None and B.vdim == 1)
                      or args.workload == "bfs"):
        def allgather_bytes(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        pb = gb.PeerBuf(A.vlen if args.workload == "bfs" else A.vdim, w["semiring"].ztype, rank, world,
                        allgather_bytes)
        exchange = {"pb": pb, "bytes": 0}

    def step_device():
        out = {"flops": 0, "nnz": 0, "device_ms": 0.0, "kernel_ms": 0.0, "nvec": 0, "infos": []}
        for (_, _, _, dm, da, db) in calls:
            if exchange is not None:
                # the rank's block of w goes straight into every GPU's dense copy over NVLink peer
                # memory (gb200_peerbuf_publish: one kernel on the library's stream), then a device-side
                # wait for all ranks' flags: no NCCL call, no host synchronisation inside the step
                rh, info = gb.axb_device_keep(dm, w["mask_comp"], da, db, w["semiring"], w["do_adotb"])
                exchange["pb"].publish(rh)
                gb.free_result(rh)
                exchange["pb"].wait()
                exchange["bytes"] = info["nnz"] * (np.dtype(gb.TYPES[w["semiring"].ztype][1]).itemsize + 1) * world
            elif slabs is not None:
                rh, info = gb.axb_device_keep(dm, w["mask_comp"], da, db, w["semiring"], w["do_adotb"])
                out.setdefault("checksums", []).append(float(gb.result_reduce(rh, w["semiring"].ztype, "PLUS")))
                gb.free_result(rh)                      # the slab of C is discarded
            else:
                info = gb.axb_device(dm, w["mask_comp"], da, db, w["semiring"], w["do_adotb"],
                                     fetch=False).info
            for k in ("flops", "nnz", "device_ms", "kernel_ms"):
                out[k] += info[k]
            out["nvec"] += info["nvec"] + 1
            out["infos"].append(info)
        return out

    def upload_gathered(m):
        """N > 1: a replicated operand crosses PCIe once in total -- every rank uploads 1/N of each
        array over its own link and the pieces are all-gathered over NVLink -- and the
        library takes the arrays from HBM (gb200_upload_from_device)."""
        ptrs, keep = [], []
        for arr in (m.p, m.h, m.i, m.x):
            if arr is None or arr.size == 0:
                ptrs.append(0)
                continue
            raw = arr.view(np.uint8).reshape(-1)
            chunk = -(-raw.size // world)
            chunk += (-chunk) % 16
            lo, hi = min(rank * chunk, raw.size), min((rank + 1) * chunk, raw.size)
            local = torch.empty(chunk, dtype=torch.uint8, device=device)
            if hi > lo:
                local[: hi - lo].copy_(torch.from_numpy(raw[lo:hi]), non_blocking=True)
            full = torch.empty(chunk * world, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(full, local)
            keep.append(full)
            ptrs.append(full.data_ptr())
        torch.cuda.synchronize()
        d = gb.DMatrix.from_device(m, ptrs[0], ptrs[1], ptrs[2], ptrs[3])
        del keep                                    # the library holds its own (narrowed) copy
        return d

    def step_host():
        """one step through the host entry points: host operands in, host T out.  The matrix A of a
        BFS step is uploaded once per step and shared by its level multiplies."""
        outs = []
        if world > 1 and args.workload != "bfs":
            for (m, a, b), (mfull, afull, bfull) in zip(hcalls, hfull):
                # replicated operands are gathered; the rank's own slice goes straight over PCIe
                cache = {}

                def up(x, replicated):
                    if x is None:
                        return None
                    if id(x) not in cache:
                        cache[id(x)] = upload_gathered(x) if replicated else gb.DMatrix(x)
                    return cache[id(x)]
                da_ = up(a, sliced_name != "A")
                db_ = up(b, sliced_name != "B")
                dm_ = up(m, sliced_name != "M")
                outs.append(gb.axb_device(dm_, w["mask_comp"], da_, db_, w["semiring"], w["do_adotb"],
                                          fetch=True, pinned=True))
                for dx in cache.values():
                    dx.free()
            return outs
        if slabs is not None:
            # A crosses PCIe once per step, every slab of B once; what comes back is the checksum of
            # every slab of C (the slab itself is discarded on the device)
            da = gb.DMatrix(hcalls[0][1])
            for (m, a, b) in hcalls:
                db_ = gb.DMatrix(b)
                rh, info = gb.axb_device_keep(None, False, da, db_, w["semiring"], w["do_adotb"])
                outs.append((float(gb.result_reduce(rh, w["semiring"].ztype, "PLUS")), info["nnz"]))
                gb.free_result(rh)
                db_.free()
            da.free()
            return outs
        if args.workload == "bfs":
            da = gb.DMatrix(hcalls[0][1])
            for (m, a, b) in hcalls:
                outs.append(gb.axb_device(gb.DMatrix(m), w["mask_comp"], da, gb.DMatrix(b),
                                          w["semiring"], w["do_adotb"], fetch=True, pinned=True))
            da.free()
        else:
            for (m, a, b) in hcalls:
                outs.append(gb.axb_host(m, w["mask_comp"], a, b, w["semiring"], w["do_adotb"],
                                        fetch=True, pinned=True))
        return outs

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        r = step_device()
    launches0 = gb.kernel_launches()
    # NVML is initialised BEFORE the barrier: eight processes doing it at once serialise in the driver,
    # and a rank that enters the timed region 30 ms late makes every other rank of a step with an
    # exchange wait for it (measured: SSSP at N = 8, rank times 3.4 ... 0.4 ms per step)
    sampler = ClockSampler(local_rank)
    barrier()
    with sampler as clk:
        t0 = time.perf_counter()
        gb.timer_mark(0)
        dev_ms, ker_ms = [], []
        for _ in range(args.steps):
            r = step_device()
            dev_ms.append(r["device_ms"])
            ker_ms.append(r["kernel_ms"])
        gb.timer_mark(1)
        t_events = gb.timer_elapsed_ms(0, 1) * 1e-3     # synchronises on the second mark
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t0
    launches = gb.kernel_launches() - launches0
    if exchange is not None and args.workload == "bfs":
        # outside the timed region: every rank's dense copy of every level's w equals the next frontier
        # of the single-GPU level loop
        for lv, (_, _, _, dm, da, db) in enumerate(calls):
            rh, info = gb.axb_device_keep(dm, w["mask_comp"], da, db, w["semiring"], w["do_adotb"])
            if not info["mask_applied"] and info["nnz"] > 0:    # an empty frontier slice returns early
                raise SystemExit("BFS push on N GPUs expects the fused <!v> (mask_applied)")
            exchange["pb"].publish(rh)
            gb.free_result(rh)
            exchange["pb"].wait()
            _, pres = exchange["pb"].read()
            want = levels[lv + 1][0].i if lv + 1 < len(levels) else np.zeros(0, dtype=np.int64)
            if not np.array_equal(np.nonzero(pres)[0], np.sort(want)):
                sys.stderr.write(f"BFS exchange check FAILED on rank {rank}, level {lv}\n")
                sys.exit(3)
        w["exchange_check"] = f"every rank's dense copy of w equals the next frontier at all {len(calls)} levels"
    tt = torch.tensor([t_events, t_wall], dtype=torch.float64, device=device)
    rank_ms = [t_events / args.steps * 1e3]
    if world > 1:
        allt = torch.zeros(world, dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(allt, tt[:1].clone())
        rank_ms = [float(v) / args.steps * 1e3 for v in allt.tolist()]
    fl = torch.tensor([r["flops"], r["nnz"]], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(fl, op=dist.ReduceOp.SUM)
    barrier()
    t_step = tt[0].item() / args.steps
    madds, cnz = int(fl[0].item()), int(fl[1].item())
    gflops = 2.0 * madds / t_step / 1e9

    # ---- algorithmic bytes of one step, from the operands and T's counts -------
    nbytes = lambda x: x.p.nbytes + x.i.nbytes + x.x.nbytes + (x.h.nbytes if x.h is not None else 0)
    zsize = np.dtype(gb.TYPES[w["semiring"].ztype][1]).itemsize
    ab = 0
    for (m, a, b, _, _, _), info in zip(calls, r["infos"]):
        uniq = list({id(x): x for x in (m, a, b) if x is not None}.values())
        # a complemented mask is not read by the saxpy method (GB_AxB_sequential.c:76-81)
        read = [x for x in uniq if not (x is m and w["mask_comp"] and not info["mask_applied"])]
        if args.workload == "bfs" and not w["do_adotb"]:
            # vector push: only the vectors of A named by the frontier are traversed
            lens = np.diff(a.p)[b.i]
            ab += 8 * len(a.p) + int(lens.sum()) * (8 + a.x.dtype.itemsize) \
                + b.nnz * (8 + b.x.dtype.itemsize) + info["nnz"] * (8 + zsize)
            ab += sum(x.nnz * (8 + x.x.dtype.itemsize) for x in read if x is m)
        else:
            ab += algo_bytes(read, info["nvec"], info["nnz"], zsize)

    if slabs is not None:
        # every operand counted once for the whole multiply, whatever the number of slabs
        Bw = slabs["whole_B"]
        ab = 8 * (len(A.p) + len(Bw.p) + Bw.nvec + 1) + A.nnz * (8 + A.x.dtype.itemsize) \
            + Bw.nnz * (8 + Bw.x.dtype.itemsize) + int(r["nnz"]) * (8 + zsize)
        slabs["checksums"] = r.get("checksums")
        slabs["checksum_total"] = float(np.sum(r.get("checksums", [0.0])))

    # ---- e2e: host buffers in, host T out, every step -----------------------------------------
    t_e2e, h2d, d2h, brk = None, 0, 0, None
    if not args.no_e2e:
        e2e_steps = 1 if slabs is not None else max(1, min(args.steps, 3))
        rh = None
        step_host()
        if slabs is None:
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            rh = None               # T of the previous step goes back to the host allocator first
            rh = step_host()
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        t_e2e = te[0].item() / e2e_steps
        seen_once = set()
        if slabs is not None:
            h2d = nbytes(hcalls[0][1]) + sum(nbytes(b) for (_, _, b) in hcalls)
            d2h = 8 * len(hcalls)
            rh = []
        for (m, a, b), res in zip(hcalls, rh):
            mine_x = {"M": m, "A": a, "B": b}.get(sliced_name)
            for x in {id(x): x for x in (m, a, b) if x is not None}.values():
                if args.workload == "bfs" and x is a:
                    if id(x) in seen_once:
                        continue                    # A of a BFS step crosses PCIe once per step
                    seen_once.add(id(x))
                # N > 1: a rank uploads its own slice whole and 1/N of every replicated operand
                h2d += nbytes(x) if (world == 1 or x is mine_x) else -(-nbytes(x) // world)
            d2h += nbytes(res.matrix)
        rh = res = None
        # where one end-to-end step spends its time: the same step through the three entry points
        # gb200_AxB_host is made of (upload, multiply on resident operands, fetch), each synchronous
        brk = {"upload_ms": 0.0, "multiply_ms": 0.0, "fetch_ms": 0.0}
        if args.workload != "bfs" and world == 1 and slabs is None:
            import ctypes as _C
            for (m, a, b) in hcalls:
                t1 = time.perf_counter()
                ha = gb.DMatrix(a)
                hb = ha if b is a else gb.DMatrix(b)
                hm = None if m is None else (ha if m is a else (hb if m is b else gb.DMatrix(m)))
                t2 = time.perf_counter()
                rhd = _C.c_void_p()
                sc = w["semiring"].c()
                gb._check(gb.lib.gb200_AxB_device(_C.byref(rhd), hm._h if hm is not None else None,
                                                  int(w["mask_comp"]), ha._h, hb._h, _C.byref(sc),
                                                  int(w["do_adotb"]), 0), "gb200_AxB_device")
                t3 = time.perf_counter()
                gb._fetch(rhd, True, True)
                t4 = time.perf_counter()
                for hx in {id(x): x for x in (ha, hb, hm) if x is not None}.values():
                    hx.free()
                brk["upload_ms"] += (t2 - t1) * 1e3
                brk["multiply_ms"] += (t3 - t2) * 1e3
                brk["fetch_ms"] += (t4 - t3) * 1e3
    pinned.clear()
    hcalls = hfull = []

    # ---- t_api: the unmodified reference's GrB_mxm / mxv / vxm with the shim interposed --------
    api = None
    if slabs is not None:
        api = {"ms": None, "skipped": "the unmodified GrB_mxm needs the whole T in host memory; the slab-"
                                      "streamed multiply is a caller of the C ABI"}
    elif G is not None and world == 1 and rank == 0 and not args.no_api:
        try:
            before = G.shim_stats()
            G.use_gpu(True)
            api_run(G, args.workload, w, 1, 1, keep=False, reps=1)          # warm-up call
            res = api_run(G, args.workload, w, 1, 1, keep=False, reps=2)
            after = G.shim_stats()
            api = {"ms": res["seconds"] * 1e3, "value": 2.0 * madds / res["seconds"] / 1e9,
                   "unit": "GFLOP/s",
                   "call": {"tri": "GrB_mxm + GrB_reduce (Demo/Source/tricount.c:166-178)", "spgemm": "GrB_mxm", "spgemm_rmat": "GrB_mxm",
                            "sssp": "GrB_mxv", "bfs": "GrB_vxm x levels"}[args.workload],
                   "through": "unmodified reference library (oracle/_ref) + libgb_b200_shim.so "
                              "(GB_AxB_parallel interposed), GxB_init with gb200_host_* (page-locked "
                              "GraphBLAS arrays); wall clock of the call, best of 2",
                   "gpu_calls": after["gpu_calls"] - before["gpu_calls"],
                   "declined": after["declined"] - before["declined"]}
            # the same call with operand residency on (gb200_cache_*): the matrix operands stay in HBM
            # between calls, which is what a BFS / SSSP / k-truss loop over one graph sees
            G.shim_cache(True)
            api_run(G, args.workload, w, 1, 1, keep=False, reps=1)          # fills the cache
            c0 = G.shim_cache()
            res = api_run(G, args.workload, w, 1, 1, keep=False, reps=2)
            c1 = G.shim_cache(False)
            api["resident"] = {"ms": res["seconds"] * 1e3, "value": 2.0 * madds / res["seconds"] / 1e9,
                               "cache_hits": c1["hits"] - c0["hits"],
                               "cache_misses": c1["misses"] - c0["misses"],
                               "note": "operand residency cache on (off in every other leg; e2e "
                                       "uploads its operands every step)"}
        except Exception as e:
            api = {**(api or {}), "error": repr(e)}
        finally:
            G.use_gpu(False)
            try:
                G.shim_cache(False)
            except Exception:
                pass

    # ---- roofline of the dominant (semiring) kernels, rank 0's slice --------------------------
    peak, peak_src = measured_peak()
    k_ms = float(np.mean(ker_ms))
    achieved = ab / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = traffic_note = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            key = args.workload + ("_pull" if args.workload == "bfs" and w["do_adotb"] else "")
            tj = json.load(f)
            traffic = tj.get(f"{key}_s{args.scale}")
            traffic_note = tj.get(f"{key}_s{args.scale}_note")
    except Exception:
        pass

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": gflops, "unit": "GFLOP/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
                "config": w["config"],
                "detail": {"madds_per_step": madds, "nnz_T": cnz, "multiplies_per_step": len(calls),
                           "timing": "CUDA events on the library's launching stream around the K "
                                     "steps, max over ranks; wall clock alongside",
                           "partition": "none (one GPU)" if world == 1 else w["partition"] if "partition" in w else
                                        (f"{world} parts of the mask's entries by owner vector (B(:,j) with j in "
                                         "the rank's range, or A(:,i) with i in it), balanced by walk length"
                                         if sliced_name == "M" else
                                         f"{world} flop-balanced contiguous slices of "
                                         f"{ {'B': 'B', 'A': 'A', 'none': 'nothing'}[sliced_name]}'s vectors"),
                           "calibration_ms": w.get("calibration_ms"),
                           "exchange_check": w.get("exchange_check"),
                           "slabs": None if slabs is None else {k: v for k, v in slabs.items() if k != "whole_B"},
                           "exchange": ("in-library: peer stores of the rank's block of w into every GPU's "
                                        "dense copy over NVLink (gb200_peerbuf_publish / _wait), "
                                        f"{exchange['bytes']} B stored per rank per step") if exchange else
                                       "none (independent output vectors; scalars all-reduced)"},
                "parity": parity,
                "wall_ms_per_step": tt[1].item() / args.steps * 1e3,
                "rank_ms_per_step": [round(v, 3) for v in rank_ms],
                "device_ms_per_step": float(np.mean(dev_ms)),
                "device_ms_steps": [round(float(v), 3) for v in dev_ms],
                "clocks": clk.summary(),
                "e2e": None if t_e2e is None else
                       {"value": 2.0 * madds / t_e2e / 1e9, "unit": "GFLOP/s",
                        "ms_per_step": t_e2e * 1e3, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h),
                        "host_memory": "page-locked (gb200_host_malloc)", "breakdown": brk,
                        "upload": "1/N of every replicated operand per rank over PCIe + NCCL "
                                  "all-gather over NVLink" if world > 1 else "host to device over PCIe"},
                "t_api": api,
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic,
                             **({"traffic_note": traffic_note} if traffic_note else {}),
                             **({"traffic_over_algorithmic": traffic / ab} if traffic else {}),
                             "kernel": w.get("kernel", "dot kernels" if w["do_adotb"] else
                                             "saxpy_hash_kernel / saxpy_light_kernel / saxpy_heavy_kernel"),
                             "kernel_ms": k_ms, "algorithmic_bytes": int(ab), "peak_source": peak_src,
                             "bytes_per_madd": ab / max(r["flops"], 1),
                             "step_algo_gbs": ab / (float(np.mean(dev_ms)) * 1e-3) / 1e9}}
        if cpu is not None:
            line["cpu_baseline"] = cpu
    # release the resident operands before the next workload
    for c in calls:
        for dx in c[3:]:
            if dx is not None:
                dx.free()
    del calls
    torch.cuda.empty_cache()
    return line


# the other configs of BASELINE.json, one GPU, after the headline line
SECONDARY = [
    # unmasked C=A*A (cfg 5, one GPU's worth).  Scale 16: nnz(C) = 352 M.  Scale 18 also fits (nnz(C) =
    # 2.9 G, 47 GB; measured line in profiles/r2) but fetching its T to the host takes 25 s per step.
    {"workload": "spgemm_rmat", "scale": 16, "ef": 16},
    {"workload": "sssp", "scale": 22, "ef": 16},             # cfg 4
    {"workload": "bfs", "scale": 22, "ef": 16},              # cfg 3
]


# ---------------------------------------------------------------------------------------------
# the reference arm: the compiled reference on the box's host cores, nothing of ours mapped
# ---------------------------------------------------------------------------------------------
def neighbour_lines(args):
    """SURVEY.md 8f rows f2 and f1 beside the path, each measured by its own tool in a process of its own (so
    that nothing it does can take the headline line down): tools/transpose_bench.py (C = A' of L = tril (A,-1) of
    the headline graph on the device: time, roofline, L' == U and (L')' == L bit for bit, oracle parity at scale
    14, the reference's GB_transpose on the host at scale 20) and tools/accum_mask_bench.py (C<M> = accum (C,T):
    three cases on resident operands, results checked, oracle parity, the reference's GB_accum_mask on the host)"""
    out = []
    for tool, metric in (("transpose_bench.py", "GB_transpose on the device"),
                         ("accum_mask_bench.py", "GB_accum_mask on the device")):
        cmd = [sys.executable, os.path.join(ROOT, "tools", tool), "--scale", str(args.scale),
               "--ef", str(args.ef), "--check-scale", str(min(14, args.scale))]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            if r.returncode != 0:
                out.append({"metric": metric, "error": (r.stderr or r.stdout)[-400:]})
            else:
                out.append(json.loads(r.stdout.strip().splitlines()[-1]))
        except Exception as e:
            out.append({"metric": metric, "error": repr(e)})
    return out


def reference_arm(args):
    import grbref
    C = pure_containers()
    w = make_workload(C, args, "cpu")
    G = grbref.GraphBLAS.get(with_shim=False)
    T = max(1, args.cpu_threads or (os.cpu_count() or 1))
    # the madds of the whole workload are not known without running it: use the measured figure of
    # the default graph to bound the sample, and thin it further after the first step if needed
    if args.cpu_stride <= 0:
        args.cpu_stride = {"tri": 4, "spgemm": 1, "spgemm_rmat": 8, "sssp": 1, "bfs": 1}[args.workload]
    vals = []
    for s in range(args.warmup + args.steps):
        res = api_run(G, args.workload, w, args.cpu_stride, T)
        if s >= args.warmup:
            vals.append((res["gflops"], res["seconds"]))
        if s == 0 and args.workload not in ("bfs",):
            # keep the whole run within a few minutes whatever K and W are: thin the sample
            total = res["seconds"] * (args.warmup + args.steps)
            if total > 150.0:
                args.cpu_stride *= int(np.ceil(total / 150.0))
    gf = float(np.mean([v[0] for v in vals]))
    dt = float(np.mean([v[1] for v in vals]))
    line = {"impl": "reference", "metric": METRIC, "value": gf,
            "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
            "config": w["config"],
            "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": res["threads"],
                             "kind": "reference", "sample": res["desc"]},
            "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "loaded_product_library": any("libgb_b200" in ln for ln in open("/proc/self/maps"))}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="tri",
                    choices=["tri", "spgemm", "spgemm_rmat", "sssp", "bfs"])
    ap.add_argument("--bfs-dir", default="push", choices=["push", "pull"])
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--ef", type=int, default=16)
    ap.add_argument("--cpu-stride", type=int, default=0, help="sample every n-th output vector "
                    "(0: sized for about 12 s of CPU work)")
    ap.add_argument("--cpu-threads", type=int, default=0, help="0: all host cores")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the parity gate")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-api", action="store_true", help="skip the t_api leg")
    ap.add_argument("--no-secondary", action="store_true", help="only the headline workload")
    ap.add_argument("--no-neighbours", action="store_true", help="skip the lines of the path's neighbours "
                    "(the device transpose, SURVEY.md 8f row f2)")
    ap.add_argument("--secondary-scale", type=int, default=0, help="run the secondary workloads at "
                    "this scale instead of their own (quick checks)")
    ap.add_argument("--slab-gb", type=float, default=0.0, help="unmasked saxpy: budget (GiB) of one slab "
                    "of C; 0: 50 %% of the free HBM")
    ap.add_argument("--slice-of", type=int, default=0, help="one GPU: run the slice --slice-rank of a "
                    "K-way partition (what one rank of --gpus K computes), for profiling")
    ap.add_argument("--slice-rank", type=int, default=0)
    ap.add_argument("--calibrate", type=int, default=2, help="N > 1, masked dot: rounds of measured "
                    "re-balancing of the ranks' parts before the timed region")
    args = ap.parse_args()
    explicit = " ".join(sys.argv[1:])
    if args.workload == "spgemm" and "--scale" not in explicit:
        args.scale, args.ef = 20, 8
    if args.workload == "spgemm_rmat" and "--scale" not in explicit:
        args.scale = 16

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # ---- reference arm: rank 0 only, CPU, no product library ------------------------------------
    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return

    # ---- our arm ----------------------------------------------------------------------------
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU path); use --impl reference for "
                         "the CPU arm")
    import torch.distributed as dist
    device = f"cuda:{local_rank}"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    os.environ["GB200_DEVICE"] = str(local_rank)
    import graphblas_b200 as gb
    gb.init(local_rank)
    env = {"rank": rank, "world": world, "local_rank": local_rank, "device": device}

    line = measure(args, env)
    default_line = (args.workload == "tri" and args.slice_of <= 1)
    if world == 1 and default_line and not args.no_secondary:
        sec = []
        for cfg in SECONDARY:
            a2 = copy.copy(args)
            a2.workload, a2.scale, a2.ef = cfg["workload"], cfg["scale"], cfg["ef"]
            if args.secondary_scale > 0:
                a2.scale = min(a2.scale, args.secondary_scale)
            a2.cpu_stride = 0
            a2.steps, a2.warmup = max(3, min(args.steps, 5)), 3
            try:
                l2 = measure(a2, env, primary=False)
                keep = ("value", "unit", "ms_per_step", "dtype", "config", "detail", "parity", "e2e",
                        "t_api", "roofline", "cpu_baseline", "gpu_launches", "device_ms_per_step")
                sec.append({k: l2[k] for k in keep if k in l2})
            except SystemExit:
                raise
            except Exception as e:      # a secondary line must not take the headline down
                sec.append({"config": cfg, "error": repr(e)})
        line["secondary"] = sec
        if not args.no_neighbours:
            line["neighbours"] = neighbour_lines(args)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
