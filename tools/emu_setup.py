"""tools/emu_setup.py -- host emulation of the masked dot's set-up kernels (no GPU needed).

The set-up kernels of engine_dot.cu (classification with the trim, pair lists, task records; and the
leaner classify -> scan -> scatter variant, GB200_DOTG_SETUP=2) are per-element loops without
shared memory.  This script cuts their SOURCE TEXT out of the .cu/.cuh files, compiles it with g++
behind a few macros (a grid is two nested loops, atomics are plain adds, shuffles return nothing), and
checks on random CSC inputs (standard and hypersparse) that
  * every task is a piece of the walked list of its pair, the pieces of a pair tile exactly the part of
    the walked list inside [first, last] of the owner, and only pairs with an empty trimmed walk are dropped;
  * both pipelines give every owner the same multiset of tasks and the same list of small pairs.

    python tools/emu_setup.py [--cases 200]
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "graphblas_b200", "csrc")


def cut(text: str, start: str, stop_after_brace: bool = True) -> str:
    """the definition that starts at `start` up to its closing brace at column 0"""
    a = text.index(start)
    m = re.compile(r"^\}", re.M).search(text, a)
    return text[a:m.end()] + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--keep", action="store_true")
    args = ap.parse_args()
    common = open(os.path.join(CS, "common.cuh")).read()
    kern = open(os.path.join(CS, "kernels.cuh")).read()
    eng = open(os.path.join(CS, "engine_dot.cu")).read()
    have2 = "dotg_classify2_kernel" in eng
    parts = [
        cut(common, "struct DMat"),
        ";\n",
        cut(common, "__device__ __forceinline__ int64_t dm_vecname"),
        "constexpr int DOTG_SEG = %s ;\nconstexpr int DOTG_SMALL = %s ;\n" % (
            re.search(r"constexpr int DOTG_SEG = (\d+)", kern).group(1),
            re.search(r"constexpr int DOTG_SMALL = (\d+)", kern).group(1)),
        "struct DotTask { int32_t e ; int32_t len ; int64_t w0 ; } ;\n",
        cut(kern, "__device__ __forceinline__ int64_t dm_vecpos"),
        cut(kern, "__device__ __forceinline__ bool dot_walkA"),
        cut(eng, "__device__ __forceinline__ void dotg_trim"),
        cut(eng, "__global__ void expand_vec_kernel"),
        cut(eng, "__global__ void dot_cum_list_kernel"),
        cut(eng, "__global__ void dotg_classify_kernel"),
        cut(eng, "__global__ void dotg_lists_kernel"),
        cut(eng, "__global__ void dotg_tasks_kernel"),
    ]
    if have2:
        parts += [
            "constexpr int DOTG_PK_SHIFT = 33 ;\nconstexpr int64_t DOTG_PK_MASK = (1LL << DOTG_PK_SHIFT) - 1 ;\n",
            cut(eng, "__global__ void dotg_classify2_kernel"),
            cut(eng, "__global__ void dotg_scatter_kernel"),
            cut(eng, "__global__ void dotg_otoff_kernel"),
        ]
    src = HARNESS.replace("@KERNELS@", "".join(parts)).replace("@HAVE2@", "1" if have2 else "0")
    d = tempfile.mkdtemp(prefix="emu_setup_")
    cpp = os.path.join(d, "emu.cpp")
    open(cpp, "w").write(src)
    exe = os.path.join(d, "emu")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-w", "-o", exe, cpp])
    rc = subprocess.call([exe, str(args.cases)])
    if args.keep:
        print("sources kept in", d)
    sys.exit(rc)


HARNESS = r"""
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <tuple>
#include <map>
// ---- a grid is two nested loops; one thread at a time ------------------------------------------
struct Dim { int64_t x ; } ;
static Dim blockIdx, threadIdx, blockDim, gridDim ;
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
template <class T> static inline T __ldg (const T *p) { return *p ; }
static inline unsigned long long atomicAdd (unsigned long long *p, unsigned long long v) { unsigned long long o = *p ; *p += v ; return o ; }
static inline unsigned __ballot_sync (unsigned, bool p) { return p ? 1u : 0u ; }        // a warp of one lane
template <class T> static inline T __shfl_sync (unsigned, T v, int) { return v ; }
template <class T> static inline T __shfl_down_sync (unsigned, T, int) { return T (0) ; }
static inline int __ffs (unsigned v) { return v ? 1 : 0 ; }
// every launch is ONE thread (grid 1, block 1): the kernels are grid-stride loops, so that thread visits
// every element, and their warp-level code sees a warp of one lane
#define LAUNCH(grid, block, call) do { gridDim.x = (grid) ; blockDim.x = (block) ; \
    for (blockIdx.x = 0 ; blockIdx.x < gridDim.x ; blockIdx.x++) \
        for (threadIdx.x = 0 ; threadIdx.x < blockDim.x ; threadIdx.x++) { call ; } } while (0)
namespace gb200 {
@KERNELS@
}
using namespace gb200 ;

struct Host { int64_t vlen, vdim ; std::vector<int64_t> p, h ; std::vector<int32_t> i ; bool hyper ; } ;
static Host make (int64_t vlen, int64_t vdim, double density, bool hyper, int hubs)
{
    Host m ; m.vlen = vlen ; m.vdim = vdim ; m.hyper = hyper ; m.p.push_back (0) ;
    for (int64_t v = 0 ; v < vdim ; v++)
    {
        double dv = density ;
        if (hubs && (rand () % 17) == 0) dv = 0.9 ;             // a long vector
        if ((rand () % 11) == 0) dv = 0 ;                        // an empty one
        std::vector<int32_t> col ;
        for (int64_t r = 0 ; r < vlen ; r++) if ((rand () / (double) RAND_MAX) < dv) col.push_back ((int32_t) r) ;
        if (hyper && col.empty ()) continue ;
        if (hyper) m.h.push_back (v) ;
        m.i.insert (m.i.end (), col.begin (), col.end ()) ;
        m.p.push_back ((int64_t) m.i.size ()) ;
    }
    for (int q = 0 ; q < 8 ; q++) m.i.push_back (0) ;           // slack, as the upload leaves
    return m ;
}
static DMat view (const Host &m)
{
    DMat d ; memset (&d, 0, sizeof (d)) ;
    d.p = m.p.data () ; d.h = m.hyper ? m.h.data () : nullptr ; d.i = m.i.data () ; d.x = nullptr ;
    d.vlen = m.vlen ; d.vdim = m.vdim ; d.nvec = (int64_t) m.p.size () - 1 ; d.nnz = m.p.back () ;
    d.hyper = (m.hyper && d.nvec < m.vdim) ? 1 : 0 ;
    if (m.hyper && !d.hyper) d.h = nullptr ;
    return d ;
}
static std::vector<int64_t> scan (const std::vector<int64_t> &in)
{
    std::vector<int64_t> out (in.size () + 1, 0) ;
    for (size_t q = 0 ; q < in.size () ; q++) out [q+1] = (int64_t) ((uint64_t) out [q] + (uint64_t) in [q]) ;
    return out ;
}
typedef std::tuple<int32_t, int32_t, int64_t> Tk ;     // e, len, w0
typedef std::map<int64_t, std::vector<Tk>> ByOwner ;

int main (int argc, char **argv)
{
    const int ncases = (argc > 1) ? atoi (argv [1]) : 100 ;
    srand (12345) ;
    long bad = 0, pairs_seen = 0, tasks_seen = 0, trimmed_away = 0 ;
    for (int cs = 0 ; cs < ncases && bad == 0 ; cs++)
    {
        const int64_t n = 40 + rand () % 2500 ;                 // vlen of A and B: long enough for split walks
        const int64_t na = 20 + rand () % 60, nb = 20 + rand () % 60 ;
        const bool hy = (cs % 3) == 1 ;
        Host Ah = make (n, na, 0.02 + 0.1 * (rand () % 5), hy, 1) ;
        Host Bh = make (n, nb, 0.02 + 0.1 * (rand () % 5), hy && (cs % 2), 1) ;
        Host Mh = make (na, nb, 0.3, (cs % 5) == 2, 0) ;
        DMat A = view (Ah), B = view (Bh), M = view (Mh) ;
        const int64_t mnz = M.nnz, anvec = A.nvec ;
        if (mnz == 0) continue ;
        const int trim = (cs % 7) != 6 ;
        std::vector<int32_t> mvec (mnz) ;
        // expand_vec_kernel is a warp kernel; do its job directly
        for (int64_t v = 0 ; v < M.nvec ; v++) for (int64_t e = M.p [v] ; e < M.p [v+1] ; e++) mvec [e] = (int32_t) v ;
        // ---------------------------------------------------------------------------------------
        // pipeline 1 (the measured one)
        std::vector<uint8_t> own (mnz), small (mnz) ;
        std::vector<int32_t> wl (mnz), ws (mnz) ;
        std::vector<unsigned long long> cntA (anvec + 1, 0), curA (anvec + 1, 0) ;
        LAUNCH (1, 1, dotg_classify_kernel (A, B, M, mvec.data (), mnz, trim, own.data (), small.data (), wl.data (), ws.data (), cntA.data ())) ;
        std::vector<int64_t> t8 (mnz) ;
        for (int64_t e = 0 ; e < mnz ; e++) t8 [e] = own [e] ;
        std::vector<int64_t> pos0 = scan (t8) ;
        for (int64_t e = 0 ; e < mnz ; e++) t8 [e] = small [e] ;
        std::vector<int64_t> poss = scan (t8) ;
        std::vector<int64_t> ca (anvec) ;
        for (int64_t v = 0 ; v < anvec ; v++) ca [v] = (int64_t) cntA [v] ;
        std::vector<int64_t> offA = scan (ca) ;
        const int64_t n0 = pos0 [mnz], ns = poss [mnz], n1 = offA [anvec] ;
        std::vector<int32_t> plist (mnz + 1), slist (mnz + 1) ;
        std::vector<int64_t> ntall (n0 + n1 + 1, 0), off0 (M.nvec + 1) ;
        LAUNCH (1, 1, dotg_lists_kernel (A, M, own.data (), small.data (), wl.data (), mnz, pos0.data (), poss.data (), offA.data (), curA.data (), n0, plist.data (), slist.data (), ntall.data ())) ;
        LAUNCH (1, 1, dot_cum_list_kernel (M.p, pos0.data (), M.nvec, off0.data ())) ;
        ByOwner own1 [2] ;
        for (int orient = 0 ; orient < 2 ; orient++)
        {
            const int64_t np = orient ? n1 : n0 ;
            if (np == 0) continue ;
            const int32_t *pl = plist.data () + (orient ? n0 : 0) ;
            std::vector<int64_t> nt (ntall.begin () + (orient ? n0 : 0), ntall.begin () + (orient ? n0 : 0) + np) ;
            std::vector<int64_t> toff = scan (nt) ;
            std::vector<DotTask> tasks (toff [np] + 1) ;
            LAUNCH (1, 1, dotg_tasks_kernel (A, B, M, mvec.data (), orient, pl, np, wl.data (), ws.data (), toff.data (), tasks.data ())) ;
            const int64_t nown = orient ? anvec : M.nvec ;
            const int64_t *off = orient ? offA.data () : off0.data () ;
            std::vector<int64_t> otoff (nown + 1) ;
            LAUNCH (1, 1, dot_cum_list_kernel (off, toff.data (), nown, otoff.data ())) ;
            for (int64_t v = 0 ; v < nown ; v++)
                for (int64_t t = otoff [v] ; t < otoff [v+1] ; t++)
                    own1 [orient][v].push_back (Tk (tasks [t].e, tasks [t].len, tasks [t].w0)) ;
            tasks_seen += toff [np] ;
        }
        std::vector<int32_t> s1 (slist.begin (), slist.begin () + ns) ;
        // ---- the tasks against first principles -----------------------------------------------
        {
            std::map<int32_t, std::vector<std::pair<int64_t,int32_t>>> pieces ;       // e -> (w0, |len|)
            for (int orient = 0 ; orient < 2 ; orient++)
                for (auto &kv : own1 [orient]) for (auto &t : kv.second)
                {
                    const int32_t e = std::get<0> (t), len = std::get<1> (t) ;
                    pieces [e].push_back (std::make_pair (std::get<2> (t), len < 0 ? -len : len)) ;
                    // the owner of the task is the vector its item says
                    const int64_t ka = dm_vecpos (A, M.i [e]), kb = dm_vecpos (B, dm_vecname (M, mvec [e])) ;
                    if ((orient ? ka : (int64_t) mvec [e]) != kv.first) { bad++ ; printf ("case %d: task of e=%d under the wrong owner\n", cs, e) ; }
                    (void) kb ;
                }
            for (int64_t e = 0 ; e < mnz ; e++)
            {
                pairs_seen++ ;
                const int64_t ka = dm_vecpos (A, M.i [e]), kb = dm_vecpos (B, dm_vecname (M, mvec [e])) ;
                int64_t want0 = 0, want1 = 0 ; bool is_small = false ;
                if (ka >= 0 && kb >= 0)
                {
                    const int64_t pa = A.p [ka], pae = A.p [ka+1], pb = B.p [kb], pbe = B.p [kb+1] ;
                    if (pae > pa && pbe > pb)
                    {
                        const bool walkA = dot_walkA (pae - pa, pbe - pb, A.vlen) ;
                        const int64_t olen = walkA ? (pbe - pb) : (pae - pa) ;
                        if (olen < DOTG_SMALL) is_small = true ;
                        else
                        {
                            const int32_t *Wi = walkA ? A.i : B.i, *Oi = walkA ? B.i : A.i ;
                            const int64_t w0 = walkA ? pa : pb, w1 = walkA ? pae : pbe ;
                            const int64_t o0 = walkA ? pb : pa, o1 = walkA ? pbe : pae ;
                            want0 = w0 ; want1 = w1 ;
                            if (trim)
                            {
                                while (want0 < w1 && Wi [want0] < Oi [o0]) want0++ ;
                                want1 = want0 ;
                                while (want1 < w1 && Wi [want1] <= Oi [o1-1]) want1++ ;
                            }
                            if (want1 - want0 < w1 - w0) trimmed_away += (w1 - w0) - (want1 - want0) ;
                        }
                    }
                }
                auto it = pieces.find ((int32_t) e) ;
                const bool in_small = std::find (s1.begin (), s1.end (), (int32_t) e) != s1.end () ;
                if (is_small != in_small) { bad++ ; printf ("case %d: small list wrong for e=%ld\n", cs, (long) e) ; }
                if (want1 <= want0) { if (it != pieces.end ()) { bad++ ; printf ("case %d: tasks for a dead pair e=%ld\n", cs, (long) e) ; } continue ; }
                if (it == pieces.end ()) { bad++ ; printf ("case %d: no task for e=%ld\n", cs, (long) e) ; continue ; }
                auto v = it->second ; std::sort (v.begin (), v.end ()) ;
                int64_t at = want0 ;
                for (auto &pc : v)
                {
                    if (pc.first != at || pc.second < 1 || pc.second > DOTG_SEG) { bad++ ; printf ("case %d: pieces of e=%ld do not tile\n", cs, (long) e) ; break ; }
                    at += pc.second ;
                }
                if (at != want1) { bad++ ; printf ("case %d: pieces of e=%ld end at %ld, want %ld\n", cs, (long) e, (long) at, (long) want1) ; }
                const bool split = v.size () > 1 ;
                for (auto &t : own1 [0]) for (auto &q : t.second) if (std::get<0> (q) == e && ((std::get<1> (q) < 0) != split)) { bad++ ; printf ("case %d: split flag\n", cs) ; }
                for (auto &t : own1 [1]) for (auto &q : t.second) if (std::get<0> (q) == e && ((std::get<1> (q) < 0) != split)) { bad++ ; printf ("case %d: split flag\n", cs) ; }
            }
        }
#if @HAVE2@
        // ---------------------------------------------------------------------------------------
        // pipeline 2 (classify -> one scan -> scatter)
        {
            std::vector<uint8_t> cls (mnz) ;
            std::vector<int32_t> wl2 (mnz) ;
            std::vector<int64_t> w0abs (mnz), pk (mnz) ;
            std::vector<unsigned long long> cntA2 (anvec + 1, 0), curA2 (anvec + 1, 0) ;
            unsigned long long tot0 = 0 ;
            LAUNCH (1, 1, dotg_classify2_kernel (A, B, M, mvec.data (), mnz, trim, cls.data (), wl2.data (), w0abs.data (), pk.data (), cntA2.data (), &tot0)) ;
            std::vector<int64_t> pos2 = scan (pk) ;
            std::vector<int64_t> ca2 (anvec) ;
            for (int64_t v = 0 ; v < anvec ; v++) ca2 [v] = (int64_t) cntA2 [v] ;
            std::vector<int64_t> offA2 = scan (ca2) ;
            const int64_t nt0 = pos2 [mnz] & DOTG_PK_MASK, ns2 = (int64_t) (((uint64_t) pos2 [mnz]) >> DOTG_PK_SHIFT), nt1 = offA2 [anvec] ;
            if ((int64_t) tot0 != nt0) { bad++ ; printf ("case %d: exact task count %llu != packed %ld\n", cs, tot0, (long) nt0) ; }
            std::vector<DotTask> tk0 (nt0 + 1), tk1 (nt1 + 1) ;
            std::vector<int32_t> slist2 (mnz + 1) ;
            std::vector<int64_t> otoff0 (M.nvec + 1) ;
            LAUNCH (1, 1, dotg_scatter_kernel (A, M, cls.data (), wl2.data (), w0abs.data (), pos2.data (), mnz, offA2.data (), curA2.data (), tk0.data (), tk1.data (), slist2.data ())) ;
            LAUNCH (1, 1, dotg_otoff_kernel (M.p, pos2.data (), M.nvec, otoff0.data ())) ;
            ByOwner own2 [2] ;
            for (int64_t v = 0 ; v < M.nvec ; v++) for (int64_t t = otoff0 [v] ; t < otoff0 [v+1] ; t++) own2 [0][v].push_back (Tk (tk0 [t].e, tk0 [t].len, tk0 [t].w0)) ;
            for (int64_t v = 0 ; v < anvec ; v++) for (int64_t t = offA2 [v] ; t < offA2 [v+1] ; t++) own2 [1][v].push_back (Tk (tk1 [t].e, tk1 [t].len, tk1 [t].w0)) ;
            if (otoff0 [M.nvec] != nt0) { bad++ ; printf ("case %d: otoff end\n", cs) ; }
            for (int orient = 0 ; orient < 2 ; orient++)
            {
                for (auto &kv : own1 [orient]) std::sort (kv.second.begin (), kv.second.end ()) ;
                for (auto &kv : own2 [orient]) std::sort (kv.second.begin (), kv.second.end ()) ;
                if (own1 [orient] != own2 [orient]) { bad++ ; printf ("case %d orient %d: the two set-ups give different tasks (%zu vs %zu owners)\n", cs, orient, own1 [orient].size (), own2 [orient].size ()) ; }
            }
            std::vector<int32_t> s2 (slist2.begin (), slist2.begin () + ns2) ;
            if (s1 != s2) { bad++ ; printf ("case %d: small lists differ (%zu vs %zu)\n", cs, s1.size (), s2.size ()) ; }
        }
#endif
    }
    printf ("emu_setup: %d cases, %ld pairs, %ld tasks, %ld walked indices trimmed away, second pipeline %s: %s\n",
        ncases, pairs_seen, tasks_seen, trimmed_away, @HAVE2@ ? "checked" : "absent", bad ? "FAILED" : "ok") ;
    return bad != 0 ;
}
"""

if __name__ == "__main__":
    main()
