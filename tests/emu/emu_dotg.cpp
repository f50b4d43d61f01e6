// tests/emu/emu_dotg.cpp -- harness of tools/emu_kernels.py: the masked dot C<M> = A'*B run on the host
// from the CUDA sources (kernels_emu.cuh = kernels.cuh without its launchers, setup_emu.cuh = the set-up
// kernels of engine_dot.cu), sequenced as run_dot (engine_dot.cu) sequences it, checked pair by pair
// against a plain intersection loop.  TEST INFRASTRUCTURE: nothing here is shipped.
#include <cuda_runtime.h>           // tests/emu/cuda_runtime.h
#include <vector>
#include <algorithm>
#include <random>
#include "kernels_emu.cuh"
#include "setup_emu.cuh"

using namespace gb200 ;

template <class T> struct Host
{
    int64_t vlen, vdim ;
    std::vector<int64_t> p, h ;
    std::vector<int32_t> i ;
    std::vector<T> x ;
    bool hyper = false, iso = false ;
} ;

// vectors of chosen lengths (0 = empty), sorted random indices; values small and non-zero, or all one
template <class T> static Host<T> make (int64_t vlen, const std::vector<int64_t> &lens, bool iso, bool hyper,
    std::mt19937 &rng)
{
    Host<T> m ; m.vlen = vlen ; m.vdim = (int64_t) lens.size () ; m.iso = iso ; m.hyper = hyper ;
    m.p.push_back (0) ;
    std::vector<int32_t> all (vlen) ;
    for (int64_t r = 0 ; r < vlen ; r++) all [r] = (int32_t) r ;
    for (int64_t v = 0 ; v < m.vdim ; v++)
    {
        const int64_t len = std::min<int64_t> (lens [v], vlen) ;
        if (hyper && len == 0) continue ;
        if (hyper) m.h.push_back (v) ;
        std::vector<int32_t> col ;
        if (len == vlen) col = all ;
        else
        {
            // a random window of the index range, so that the lists of a pair overlap only partly
            const int64_t span = std::min<int64_t> (vlen, std::max<int64_t> (len, (int64_t) (len * (1.2 + (rng () % 40) / 10.0)))) ;
            const int64_t lo = (int64_t) (rng () % (uint64_t) (vlen - span + 1)) ;
            std::vector<int32_t> cand (span) ;
            for (int64_t q = 0 ; q < span ; q++) cand [q] = (int32_t) (lo + q) ;
            std::shuffle (cand.begin (), cand.end (), rng) ;
            col.assign (cand.begin (), cand.begin () + len) ;
            std::sort (col.begin (), col.end ()) ;
        }
        for (auto r : col)
        {
            m.i.push_back (r) ;
            if (iso) m.x.push_back ((T) 1) ;
            else if constexpr (std::is_same<T, bool>::value) m.x.push_back ((rng () % 4) != 0) ;
            else if constexpr (std::is_floating_point<T>::value) m.x.push_back ((T) (0.25 * (double) (1 + rng () % 64))) ;
            else m.x.push_back ((T) (1 + rng () % 5)) ;
        }
        m.p.push_back ((int64_t) m.i.size ()) ;
    }
    for (int q = 0 ; q < 8 ; q++) m.i.push_back (0) ;            // the upload leaves 32 bytes of slack
    if (m.x.empty ()) m.x.push_back ((T) 1) ;
    return m ;
}

// std::vector<bool> has no data(): values live in a plain byte buffer
template <class T> struct Vals { std::vector<unsigned char> raw ; const T *ptr = nullptr ; } ;
template <class T> static Vals<T> values (const Host<T> &m)
{
    Vals<T> v ; v.raw.resize (m.x.size () * sizeof (T) + 16) ;
    for (size_t q = 0 ; q < m.x.size () ; q++) { T t = m.x [q] ; memcpy (v.raw.data () + q * sizeof (T), &t, sizeof (T)) ; }
    v.ptr = (const T *) v.raw.data () ;
    return v ;
}

template <class T> static DMat view (const Host<T> &m, const T *x)
{
    DMat d ; memset (&d, 0, sizeof (d)) ;
    d.p = m.p.data () ; d.i = m.i.data () ; d.x = x ;
    d.vlen = m.vlen ; d.vdim = m.vdim ; d.nvec = (int64_t) m.p.size () - 1 ; d.nnz = m.p.back () ;
    d.hyper = (m.hyper && d.nvec < m.vdim) ? 1 : 0 ;
    d.h = d.hyper ? m.h.data () : nullptr ;
    d.iso = m.iso ? 1 : 0 ;
    return d ;
}

static std::vector<int64_t> scan (const std::vector<int64_t> &in)
{
    std::vector<int64_t> out (in.size () + 1, 0) ;
    for (size_t q = 0 ; q < in.size () ; q++) out [q+1] = out [q] + in [q] ;
    return out ;
}

static long g_bad = 0 ;

// one masked dot, sequenced as run_dot does it; ISO selects the pattern-only kernels
template <class S, bool ISO> static void run_case (const char *name, int64_t n, const std::vector<int64_t> &lensA,
    const std::vector<int64_t> &lensB, double mdensity, bool hyper, int mult_op, uint32_t seed, int trim)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    std::mt19937 rng (seed) ;
    Host<T> Ah = make<T> (n, lensA, ISO, hyper, rng), Bh = make<T> (n, lensB, ISO, false, rng) ;
    // the mask: pattern-only values; every vector of A x B pair with probability mdensity
    Host<bool> Mh ; Mh.vlen = Ah.vdim ; Mh.vdim = Bh.vdim ; Mh.p.push_back (0) ;
    for (int64_t j = 0 ; j < Mh.vdim ; j++)
    {
        for (int64_t i = 0 ; i < Mh.vlen ; i++)
            if ((rng () % 1000) < (uint32_t) (mdensity * 1000)) Mh.i.push_back ((int32_t) i) ;
        Mh.p.push_back ((int64_t) Mh.i.size ()) ;
    }
    for (int q = 0 ; q < 8 ; q++) Mh.i.push_back (0) ;
    Vals<T> Ax = values (Ah), Bx = values (Bh) ;
    DMat A = view (Ah, Ax.ptr), B = view (Bh, Bx.ptr) ;
    DMat M ; memset (&M, 0, sizeof (M)) ;
    M.p = Mh.p.data () ; M.i = Mh.i.data () ; M.vlen = Mh.vlen ; M.vdim = Mh.vdim ; M.nvec = Mh.vdim ; M.nnz = Mh.p.back () ;
    const int64_t mnz = M.nnz, anvec = A.nvec ;
    if (mnz == 0) return ;
    const S sr (mult_op, false) ;
    const int64_t cap = dotg_cap (ISO) ;

    std::vector<int32_t> mvec (mnz) ;
    emu::launch (2, 64, [&] { expand_vec_kernel (M.p, M.nvec, mnz, mvec.data ()) ; }) ;
    for (int64_t v = 0 ; v < M.nvec ; v++) for (int64_t e = M.p [v] ; e < M.p [v+1] ; e++)
        if (mvec [e] != v) { g_bad++ ; printf ("%s: expand_vec wrong at %ld\n", name, (long) e) ; return ; }

    std::vector<acc_t> vals (mnz, Mon::identity ()) ;
    std::vector<uint8_t> flags (mnz, 0) ;
    unsigned long long nmatch = 0 ;

    // ---- set-up, as run_dot ---------------------------------------------------------------------
    std::vector<uint8_t> own (mnz), small (mnz) ;
    std::vector<int32_t> wl (mnz), ws (mnz) ;
    std::vector<unsigned long long> cntA (anvec + 1, 0), curA (anvec + 1, 0) ;
    emu::launch (2, 64, [&] { dotg_classify_kernel (A, B, M, mvec.data (), mnz, trim, own.data (), small.data (),
        wl.data (), ws.data (), cntA.data ()) ; }) ;
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
    emu::launch (2, 64, [&] { dotg_lists_kernel (A, M, own.data (), small.data (), wl.data (), mnz, pos0.data (),
        poss.data (), offA.data (), curA.data (), n0, plist.data (), slist.data (), ntall.data ()) ; }) ;
    emu::launch (1, 64, [&] { dot_cum_list_kernel (M.p, pos0.data (), M.nvec, off0.data ()) ; }) ;

    // ---- small pairs: dot_kernel, 4 lanes per pair ------------------------------------------------
    if (ns > 0)
    {
        DotArgs da ; memset (&da, 0, sizeof (da)) ;
        da.A = A ; da.B = B ; da.M = M ; da.mode = DOT_MASK ; da.mvec = mvec.data () ; da.plist = slist.data () ;
        da.npairs = ns ; da.vals = vals.data () ; da.flags = flags.data () ; da.G = 4 ; da.nmatch = &nmatch ;
        da.mult_op = mult_op ; da.flip = 0 ;
        emu::launch (2, 256, [&] { dot_kernel<S> (da) ; }) ;
    }
    // ---- owner-grouped pairs: dotg_kernel -------------------------------------------------------
    long ntask_all = 0, nhubitems = 0, nregitems = 0 ;
    for (int orient = 0 ; orient < 2 ; orient++)
    {
        const int64_t np = orient ? n1 : n0 ;
        if (np == 0) continue ;
        const int32_t *pl = plist.data () + (orient ? n0 : 0) ;
        const int64_t *off = orient ? offA.data () : off0.data () ;
        const int64_t nown = orient ? anvec : M.nvec ;
        std::vector<int64_t> nt (ntall.begin () + (orient ? n0 : 0), ntall.begin () + (orient ? n0 : 0) + np) ;
        std::vector<int64_t> toff = scan (nt) ;
        const int64_t ntasks = toff [np] ;
        ntask_all += ntasks ;
        std::vector<DotTask> tasks (ntasks + 1) ;
        emu::launch (2, 64, [&] { dotg_tasks_kernel (A, B, M, mvec.data (), orient, pl, np, wl.data (), ws.data (),
            toff.data (), tasks.data ()) ; }) ;
        std::vector<int64_t> otoff (nown + 1) ;
        emu::launch (1, 64, [&] { dot_cum_list_kernel (off, toff.data (), nown, otoff.data ()) ; }) ;
        for (int hub = 1 ; hub >= 0 ; hub--)
        {
            const int64_t chunk = hub ? 2048 : 64 ;         // small items: several items per owner
            std::vector<int64_t> nch (nown + 1, 0) ;
            emu::launch (1, 64, [&] { dotg_nchunks_kernel (orient ? A : B, M, orient, cap, hub, chunk, otoff.data (),
                nown, nch.data ()) ; }) ;
            nch.resize (nown) ;
            std::vector<int64_t> ioff = scan (nch) ;
            const int64_t nitems = ioff [nown] ;
            if (nitems == 0) continue ;
            (hub ? nhubitems : nregitems) += nitems ;
            std::vector<DotItem> items (nitems) ;
            emu::launch (1, 64, [&] { dotg_items_kernel (orient ? A : B, M, orient, cap, hub, chunk, otoff.data (),
                ioff.data (), nown, items.data ()) ; }) ;
            unsigned long long next_item = 0 ;
            DotGArgs ga ; memset (&ga, 0, sizeof (ga)) ;
            ga.A = A ; ga.B = B ; ga.M = M ; ga.tasks = tasks.data () ; ga.items = items.data () ; ga.nitems = nitems ;
            ga.orient = orient ; ga.vals = vals.data () ; ga.flags = flags.data () ; ga.nmatch = &nmatch ;
            ga.next_item = &next_item ; ga.mult_op = mult_op ; ga.flip = 0 ;
            if (hub) emu::launch (2, DOTG_THREADS, [&] { dotg_kernel<S, ISO, true> (ga) ; }) ;
            else     emu::launch (2, DOTG_THREADS, [&] { dotg_kernel<S, ISO, false> (ga) ; }) ;
        }
    }
    // ---- the answer by plain intersection ---------------------------------------------------------
    unsigned long long want_matches = 0 ;
    long npresent = 0 ;
    for (int64_t v = 0 ; v < M.nvec ; v++)
    {
        for (int64_t e = M.p [v] ; e < M.p [v+1] ; e++)
        {
            const int64_t ka = dm_vecpos (A, M.i [e]), kb = v ;
            acc_t c = Mon::identity () ; bool found = false ;
            unsigned long long nm = 0 ;
            if (ka >= 0)
            {
                int64_t pa = A.p [ka], pb = B.p [kb] ;
                const int64_t pae = A.p [ka+1], pbe = B.p [kb+1] ;
                while (pa < pae && pb < pbe)
                {
                    if (A.i [pa] < B.i [pb]) pa++ ;
                    else if (A.i [pa] > B.i [pb]) pb++ ;
                    else
                    {
                        const acc_t t = sr.product (Ax.ptr [pa], Bx.ptr [pb]) ;
                        c = found ? Mon::combine (c, t) : t ; found = true ; nm++ ;
                        pa++ ; pb++ ;
                    }
                }
            }
            want_matches += nm ;
            npresent += found ;
            if ((flags [e] != 0) != found)
            { g_bad++ ; if (g_bad < 10) printf ("%s: presence of pair e=%ld (i=%d, j=%ld) is %d, want %d\n", name, (long) e, M.i [e], (long) v, flags [e], (int) found) ; }
            else if (found && memcmp (&vals [e], &c, sizeof (acc_t)) != 0)
            { g_bad++ ; if (g_bad < 10) printf ("%s: value of pair e=%ld differs\n", name, (long) e) ; }
        }
    }
    // early exit on a terminal value may stop counting matches: the count is exact only without one
    if (!(Mon::has_terminal () && !ISO) && nmatch != want_matches)
    { g_bad++ ; printf ("%s: %llu matches counted, want %llu\n", name, nmatch, want_matches) ; }
    printf ("%-44s pairs %6ld (small %5ld) tasks %6ld hub items %3ld regular items %4ld present %6ld matches %9llu : %s\n",
        name, (long) mnz, (long) ns, ntask_all, nhubitems, nregitems, npresent, want_matches, g_bad ? "FAILED" : "ok") ;
}

int main (int argc, char **argv)
{
    setvbuf (stdout, nullptr, _IOLBF, 0) ;
    const int ncases = (argc > 1) ? atoi (argv [1]) : 1 ;
    // vector lengths: hubs beyond one table load (6144 pattern-only, 3072 valued), a dense vector,
    // regular owners, owners shorter than DOTG_SMALL, empty vectors; walks longer than DOTG_SEG
    const int64_t n = 14000 ;
    std::vector<int64_t> la = { 9000, 40, 0, 700, 13000, 3, 2500, 14000, 31, 32, 120, 6144, 6145, 1500, 5, 3300, 0, 64, 900, 4000 } ;
    std::vector<int64_t> lb = { 300, 11000, 20, 0, 1025, 7000, 14000, 33, 2048, 10, 5000, 77, 3073, 1, 640, 12500 } ;
    for (int cs = 0 ; cs < ncases ; cs++)
    {
        const uint32_t seed = 1000 + 17 * cs ;
        const int trim = (cs % 3) != 2 ;
        const bool hyper = (cs % 2) == 1 ;
        run_case<Semiring<int64_t, int64_t, GB200_PLUS, GB200_TIMES>, true>  ("PLUS_TIMES_INT64 pattern-only", n, la, lb, 0.5, hyper, GB200_TIMES, seed, trim) ;
        run_case<Semiring<int64_t, int64_t, GB200_PLUS, GB200_TIMES>, false> ("PLUS_TIMES_INT64 valued", n, la, lb, 0.5, hyper, GB200_TIMES, seed + 1, trim) ;
        run_case<Semiring<double, double, GB200_MIN, GB200_PLUS>, false>     ("MIN_PLUS_FP64 valued", n, la, lb, 0.4, hyper, GB200_PLUS, seed + 2, trim) ;
        run_case<Semiring<bool, bool, GB200_LOR, GB200_LAND>, true>          ("LOR_LAND_BOOL pattern-only", n, la, lb, 0.4, hyper, GB200_LAND, seed + 3, trim) ;
        run_case<Semiring<bool, bool, GB200_LOR, GB200_LAND>, false>         ("LOR_LAND_BOOL valued (terminal exit)", n, la, lb, 0.4, hyper, GB200_LAND, seed + 4, trim) ;
        if (g_bad) break ;
    }
    printf ("emu_kernels: %s\n", g_bad ? "FAILED" : "ok") ;
    return g_bad != 0 ;
}
