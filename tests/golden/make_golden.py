"""make_golden.py -- regenerates tests/golden/*.npz.  Run in the build container only (it needs
oracle/_ref built from /root/reference and, for the tri_demo fixtures, /root/reference/Demo/Matrix).

Two families of fixtures:
  seam_*.npz   inputs + the T returned by the REFERENCE's own GB_AxB_parallel (called directly on
               GrB_Matrix handles), one per method / mask / hypersparsity combination;
  tri_*.npz    adjacency patterns built from the reference's Demo/Matrix files with the rule of
               Demo/Source/read_matrix.c (zero-based tuples, self-edges dropped, A = spones(A+A')),
               with the triangle count the reference prints in Demo/Output/tri_demo.out.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import gen                      # noqa: E402
import grbref                   # noqa: E402
import graphblas_b200 as gb     # noqa: E402
from test_oracle import seam_reference   # noqa: E402

REF = "/root/reference"
# Demo/Output/tri_demo.out line numbers of the "# triangles" lines
TRI_KNOWN = {"t1": (2, 256), "bcsstk01": (160, 464), "fs_183_1": (863, 636), "west0067": (120, 1048),
             "2blocks": (0, 206)}


def put(d, pfx, m: gb.Matrix):
    d[pfx + "_vlen"], d[pfx + "_vdim"] = m.vlen, m.vdim
    d[pfx + "_p"], d[pfx + "_i"], d[pfx + "_x"], d[pfx + "_type"] = m.p, m.i, m.x, m.type
    if m.h is not None:
        d[pfx + "_h"] = m.h


def seam_case(G, name, A, B, M, comp, sr, dot, method, exact):
    ref, used, applied = seam_reference(G, M, comp, A, B, sr, dot, method)
    T = gb.Matrix(ref["vlen"], ref["vdim"], ref["p"], ref["i"], ref["x"], ref["h"], ref["type"])
    d = {"add": sr.add, "mult": sr.mult, "xytype": sr.xytype, "flipxy": int(sr.flipxy),
         "mask_comp": int(comp), "do_adotb": int(dot), "method_used": used,
         "mask_applied": int(applied), "exact": "1" if exact else "0"}
    put(d, "A", A)
    put(d, "B", B)
    put(d, "T", T)
    if M is not None:
        put(d, "M", M)
    np.savez_compressed(os.path.join(HERE, f"seam_{name}.npz"), **d)
    print(f"seam_{name}: nnz(T)={T.nnz} method {used} mask_applied {applied}")


def typecast_cases(G):
    """operands of built-in types other than the multiply operator's: the reference typecasts them
    (GB_CAST, Source/GB.h:2925-2947), including NaN -> 0 and +-Inf -> the integer range"""
    def operand(seed, r, c, nz, t):
        S = gen.er(r, c, nz, seed, grbref.NP_OF[t], lo=-6, hi=7).tocsc()
        if t in ("FP32", "FP64"):
            S.data = (S.data * 1.37).astype(grbref.NP_OF[t])
            S.data[::9] = np.nan
            S.data[1::11] = np.inf
            S.data[2::13] = -np.inf
        return gb.Matrix.from_scipy(S, t)
    M = gb.Matrix.from_scipy(gen.er(60, 55, 700, 113, np.int8, lo=0, hi=2).tocsc())
    seam_case(G, "typecast_int32_fp32_to_fp64", operand(111, 60, 50, 400, "INT32"),
              operand(112, 50, 55, 380, "FP32"), None, False, gb.Semiring("PLUS", "TIMES", "FP64"), False,
              grbref.GxB_AxB_GUSTAVSON, True)
    seam_case(G, "typecast_fp64_to_int16_masked_dot", operand(114, 50, 60, 400, "FP64"),
              operand(115, 50, 55, 380, "FP64"), M, False, gb.Semiring("MIN", "PLUS", "INT16"), True, 0, True)


def main():
    G = grbref.GraphBLAS.get(with_shim=False)
    m = lambda s, r, c, nz, dt: gb.Matrix.from_scipy(gen.er(r, c, nz, s, dt).tocsc())
    A, B = m(101, 60, 50, 400, np.float64), m(102, 50, 55, 380, np.float64)
    M = gb.Matrix.from_scipy(gen.er(60, 55, 700, 103, np.int8, lo=0, hi=2).tocsc())
    pt = gb.Semiring("PLUS", "TIMES", "FP64")
    seam_case(G, "gustavson_fp64", A, B, None, False, pt, False, grbref.GxB_AxB_GUSTAVSON, True)
    seam_case(G, "heap_fp64", A, B, None, False, pt, False, grbref.GxB_AxB_HEAP, False)
    seam_case(G, "gustavson_masked_fp64", A, B, M, False, pt, False, grbref.GxB_AxB_GUSTAVSON, True)
    seam_case(G, "gustavson_compmask_fp64", A, B, M, True, pt, False, grbref.GxB_AxB_GUSTAVSON, True)
    At = m(104, 50, 60, 400, np.float64)
    Md = gb.Matrix.from_scipy(gen.er(60, 55, 700, 105, np.int8, lo=0, hi=2).tocsc())
    seam_case(G, "dot_fp64", At, B, None, False, pt, True, 0, True)
    seam_case(G, "dot_masked_fp64", At, B, Md, False, pt, True, 0, True)
    seam_case(G, "dot_compmask_fp64", At, B, Md, True, pt, True, 0, True)
    Ai, Bi = m(106, 60, 50, 400, np.int32), m(107, 50, 55, 380, np.int32)
    seam_case(G, "hyper_minus_int32_flip", Ai.to_hyper(), Bi.to_hyper(), M.to_hyper(), False,
              gb.Semiring("PLUS", "MINUS", "INT32", True), False, 0, True)
    Ab, Bb = m(108, 70, 70, 500, np.bool_), m(109, 70, 1, 30, np.bool_)
    seam_case(G, "lor_land_bool_vector", Ab, Bb, None, False, gb.Semiring("LOR", "LAND", "BOOL"),
              False, 0, True)
    Aw = m(110, 70, 70, 600, np.float64)
    d0 = gb.Matrix(70, 1, np.array([0, 70]), np.arange(70), np.random.default_rng(1).random(70))
    seam_case(G, "min_plus_fp64_dense_vector", Aw, d0, None, False,
              gb.Semiring("MIN", "PLUS", "FP64"), True, 0, True)

    typecast_cases(G)

    # tri_demo known answers
    for name, (ntri, line) in TRI_KNOWN.items():
        rows = np.loadtxt(os.path.join(REF, "Demo", "Matrix", name), ndmin=2)
        i, j = rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64)
        n = int(max(i.max(), j.max())) + 1
        keep = i != j
        S = sp.csr_matrix((np.ones(keep.sum(), dtype=np.int64), (i[keep], j[keep])), shape=(n, n))
        S = ((S + S.T) != 0).astype(np.int64).tocsr()
        S.sort_indices()
        check = int((S @ S).multiply(S).sum() // 6)
        assert check == ntri, (name, check, ntri)
        np.savez_compressed(os.path.join(HERE, f"tri_{name}.npz"), n=n, p=S.indptr.astype(np.int64),
                            i=S.indices.astype(np.int64), ntri=ntri, tri_demo_out_line=line)
        print(f"tri_{name}: n={n} nnz={S.nnz} ntri={ntri}")


if __name__ == "__main__":
    main()
