// tests/emu/emu_vec.cpp -- harness of tools/emu_kernels.py for the vector multiplies (B is n-by-1): the
// streamed SpMV (spmv_stream_kernel, GrB_mxv / SSSP), the masked pull (dotv_kernel + dotv_long_kernel) and
// the push (saxpyv_kernel + saxpyv_long_kernel, GrB_vxm / BFS) run on the host from kernels_vec.cuh,
// sequenced as run_dotv / run_saxpyv (engine_vec.cu) sequence them, and compared with plain loops.
// Floating-point values are small integers, so that PLUS sums are exact in any order.
// TEST INFRASTRUCTURE: nothing here is shipped.
#include <cuda_runtime.h>           // tests/emu/cuda_runtime.h
#include <vector>
#include <algorithm>
#include <random>
#include "kernels_emu.cuh"
#include "setup_vec_emu.cuh"

using namespace gb200 ;

static long g_bad = 0 ;

template <class T> static T rand_value (std::mt19937 &rng)
{
    if constexpr (std::is_same<T, bool>::value) return (rng () % 4) != 0 ;
    else return (T) (1 + rng () % 8) ;
}

template <class T> struct Mat
{
    int64_t vlen = 0, vdim = 0 ;
    std::vector<int64_t> p ;
    std::vector<int32_t> i ;
    std::vector<unsigned char> xraw ;           // values as bytes (std::vector<bool> has no data())
    const T *x () const { return (const T *) xraw.data () ; }
    DMat view () const
    {
        DMat d ; memset (&d, 0, sizeof (d)) ;
        d.p = p.data () ; d.i = i.data () ; d.x = xraw.data () ; d.vlen = vlen ; d.vdim = vdim ;
        d.nvec = vdim ; d.nnz = p.back () ; d.hyper = 0 ; d.h = nullptr ;
        return d ;
    }
} ;

template <class T> static Mat<T> make (int64_t vlen, const std::vector<int64_t> &lens, std::mt19937 &rng)
{
    Mat<T> m ; m.vlen = vlen ; m.vdim = (int64_t) lens.size () ; m.p.push_back (0) ;
    std::vector<int32_t> cand (vlen) ;
    for (int64_t v = 0 ; v < m.vdim ; v++)
    {
        const int64_t len = std::min (lens [v], vlen) ;
        for (int64_t r = 0 ; r < vlen ; r++) cand [r] = (int32_t) r ;
        if (len < vlen) { std::shuffle (cand.begin (), cand.end (), rng) ; std::sort (cand.begin (), cand.begin () + len) ; }
        for (int64_t q = 0 ; q < len ; q++)
        {
            m.i.push_back (cand [q]) ;
            const T t = rand_value<T> (rng) ;
            const size_t at = m.xraw.size () ; m.xraw.resize (at + sizeof (T)) ; memcpy (m.xraw.data () + at, &t, sizeof (T)) ;
        }
        m.p.push_back ((int64_t) m.i.size ()) ;
    }
    for (int q = 0 ; q < 8 ; q++) m.i.push_back (0) ;
    m.xraw.resize (m.xraw.size () + 16) ;
    return m ;
}

static std::vector<int64_t> scan (const std::vector<int64_t> &in)
{
    std::vector<int64_t> out (in.size () + 1, 0) ;
    for (size_t q = 0 ; q < in.size () ; q++) out [q+1] = out [q] + in [q] ;
    return out ;
}

static bool bit (const std::vector<uint32_t> &b, int64_t k) { return (b [k >> 5] >> (k & 31)) & 1u ; }

// ---- pull: w<M> = A'*u ----------------------------------------------------------------------------
// mode 0: streamed SpMV, no mask; 1: dotv kernels with a mask; 2: dotv kernels with a complemented mask
template <class S> static void run_pull (const char *name, const Mat<typename S::T> &Am, bool dense_u, int mode,
    int mult_op, uint32_t seed)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    std::mt19937 rng (seed) ;
    const DMat A = Am.view () ;
    const int64_t n = A.vlen, anvec = A.nvec ;
    const S sr (mult_op, false) ;
    std::vector<unsigned char> uraw (n * sizeof (T) + 16) ;
    T *u = (T *) uraw.data () ;
    std::vector<uint32_t> upres ((n + 31) / 32 + 1, 0), mbits ((anvec + 31) / 32 + 1, 0) ;
    for (int64_t k = 0 ; k < n ; k++)
    {
        u [k] = rand_value<T> (rng) ;
        if (dense_u || (rng () % 10) < 3) upres [k >> 5] |= (1u << (k & 31)) ;
    }
    for (int64_t v = 0 ; v < anvec ; v++) if (rng () % 2) mbits [v >> 5] |= (1u << (v & 31)) ;
    std::vector<acc_t> vals (anvec) ;
    memset (vals.data (), 0xAB, anvec * sizeof (acc_t)) ;
    std::vector<uint8_t> flags (anvec, 0xCD) ;
    unsigned long long nmatch = 0 ;
    if (mode == 0)
    {
        for (auto &v : vals) v = Mon::identity () ;
        std::fill (flags.begin (), flags.end (), 0) ;
        const int64_t ntiles = (A.nnz + SPMV_TILE - 1) / SPMV_TILE ;
        std::vector<int32_t> tile_row (ntiles + 1) ;
        emu::launch (2, 64, [&] { tile_row_kernel (A.p, anvec, ntiles, tile_row.data ()) ; }) ;
        SpmvArgs sp ; memset (&sp, 0, sizeof (sp)) ;
        sp.A = A ; sp.bval = u ; sp.bpres = dense_u ? nullptr : upres.data () ; sp.tile_row = tile_row.data () ;
        sp.ntiles = ntiles ; sp.vals = vals.data () ; sp.flags = flags.data () ; sp.nmatch = &nmatch ;
        sp.mult_op = mult_op ; sp.flip = 0 ;
        if (dense_u) emu::launch (3, SPMV_THREADS, [&] { spmv_stream_kernel<S, false, false> (sp) ; }) ;
        else         emu::launch (3, SPMV_THREADS, [&] { spmv_stream_kernel<S, true, false> (sp) ; }) ;
    }
    else
    {
        std::vector<int64_t> nseg (anvec) ;
        emu::launch (2, 64, [&] { vec_nseg_kernel (A.p, anvec, nseg.data ()) ; }) ;
        std::vector<int64_t> off = scan (nseg) ;
        std::vector<VecItem> items (off [anvec] + 1) ;
        emu::launch (2, 64, [&] { vec_items_kernel (A.p, off.data (), anvec, items.data ()) ; }) ;
        DotVArgs da ; memset (&da, 0, sizeof (da)) ;
        da.A = A ; da.bval = u ; da.bpres = dense_u ? nullptr : upres.data () ; da.mbits = mbits.data () ;
        da.mask_comp = (mode == 2) ; da.vals = vals.data () ; da.flags = flags.data () ; da.G = 8 ;
        da.items = items.data () ; da.nitems = off [anvec] ; da.nmatch = &nmatch ; da.mult_op = mult_op ; da.flip = 0 ;
        emu::launch (3, 256, [&] { dotv_kernel<S> (da) ; }) ;
        if (da.nitems > 0) emu::launch (2, 256, [&] { dotv_long_kernel<S> (da) ; }) ;
    }
    long present = 0 ;
    for (int64_t v = 0 ; v < anvec ; v++)
    {
        bool live = true ;
        if (mode == 1) live = bit (mbits, v) ;
        if (mode == 2) live = !bit (mbits, v) ;
        acc_t c = Mon::identity () ; bool found = false ;
        if (live)
            for (int64_t p = A.p [v] ; p < A.p [v+1] ; p++)
            {
                const int64_t k = A.i [p] ;
                if (!bit (upres, k)) continue ;
                const acc_t t = sr.product (Am.x () [p], u [k]) ;
                c = found ? Mon::combine (c, t) : t ; found = true ;
            }
        present += found ;
        if ((flags [v] == 1) != found || (flags [v] != 0 && flags [v] != 1))
        { g_bad++ ; if (g_bad < 10) printf ("%s: presence of w(%ld) is %d, want %d\n", name, (long) v, flags [v], (int) found) ; }
        else if (found && memcmp (&vals [v], &c, sizeof (acc_t)) != 0)
        { g_bad++ ; if (g_bad < 10) printf ("%s: value of w(%ld) differs\n", name, (long) v) ; }
    }
    printf ("%-58s vectors %4ld nnz(A) %7ld entries of w %4ld : %s\n", name, (long) anvec, (long) A.nnz, present, g_bad ? "FAILED" : "ok") ;
}

// ---- push: w<M> = A*u -----------------------------------------------------------------------------
// mode 0: no mask; 1: mask; 2: complemented mask (the BFS step)
template <class S> static void run_push (const char *name, const Mat<typename S::T> &Am, double udensity, int mode,
    int mult_op, uint32_t seed)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    std::mt19937 rng (seed) ;
    const DMat A = Am.view () ;
    const int64_t n = A.vlen, m = A.vdim ;
    const S sr (mult_op, false) ;
    // u: a sparse m-by-1 vector
    Mat<T> um ; um.vlen = m ; um.vdim = 1 ; um.p.push_back (0) ;
    for (int64_t k = 0 ; k < m ; k++)
        if ((rng () % 1000) < (uint32_t) (udensity * 1000))
        {
            um.i.push_back ((int32_t) k) ;
            const T t = rand_value<T> (rng) ;
            const size_t at = um.xraw.size () ; um.xraw.resize (at + sizeof (T)) ; memcpy (um.xraw.data () + at, &t, sizeof (T)) ;
        }
    um.p.push_back ((int64_t) um.i.size ()) ;
    if (um.i.empty ()) return ;
    um.xraw.resize (um.xraw.size () + 16) ;
    const DMat B = um.view () ;
    std::vector<uint32_t> mbits ((n + 31) / 32 + 1, 0), pres ((n + 31) / 32 + 1, 0) ;
    for (int64_t r = 0 ; r < n ; r++) if (rng () % 2) mbits [r >> 5] |= (1u << (r & 31)) ;
    std::vector<acc_t> acc (n, Mon::identity ()) ;
    std::vector<int32_t> longlist (B.nnz + 1, -1) ;
    unsigned long long counters [2] = { 0, 0 } ;
    SaxpyVArgs sa ; memset (&sa, 0, sizeof (sa)) ;
    sa.A = A ; sa.B = B ; sa.mbits = mode ? mbits.data () : nullptr ; sa.mask_comp = (mode == 2) ;
    sa.acc = acc.data () ; sa.pres = pres.data () ; sa.longlist = longlist.data () ;
    sa.hugelist = longlist.data () + (B.nnz - 1) ; sa.nlong = (unsigned int *) &counters [0] ;
    sa.nflops = &counters [1] ; sa.mult_op = mult_op ; sa.flip = 0 ;
    emu::launch (2, 256, [&] { saxpyv_kernel<S> (sa) ; }) ;
    emu::launch (3, 256, [&] { saxpyv_long_kernel<S> (sa) ; }) ;
    std::vector<acc_t> want (n, Mon::identity ()) ;
    std::vector<uint8_t> wp (n, 0) ;
    unsigned long long flops = 0 ;
    for (int64_t pb = 0 ; pb < B.nnz ; pb++)
    {
        const int64_t k = B.i [pb] ;
        flops += (unsigned long long) (A.p [k+1] - A.p [k]) ;
        for (int64_t p = A.p [k] ; p < A.p [k+1] ; p++)
        {
            const int64_t i = A.i [p] ;
            if (mode == 1 && !bit (mbits, i)) continue ;
            if (mode == 2 && bit (mbits, i)) continue ;
            const acc_t t = sr.product (Am.x () [p], um.x () [pb]) ;
            want [i] = wp [i] ? Mon::combine (want [i], t) : t ; wp [i] = 1 ;
        }
    }
    long present = 0 ;
    for (int64_t i = 0 ; i < n ; i++)
    {
        present += wp [i] ;
        if (bit (pres, i) != (wp [i] != 0))
        { g_bad++ ; if (g_bad < 10) printf ("%s: presence of w(%ld) is %d, want %d\n", name, (long) i, (int) bit (pres, i), wp [i]) ; }
        else if (wp [i] && memcmp (&acc [i], &want [i], sizeof (acc_t)) != 0)
        { g_bad++ ; if (g_bad < 10) printf ("%s: value of w(%ld) differs\n", name, (long) i) ; }
    }
    if (counters [1] != flops) { g_bad++ ; printf ("%s: %llu flops counted, want %llu\n", name, counters [1], flops) ; }
    const unsigned int *nl = (const unsigned int *) &counters [0] ;
    printf ("%-58s entries of u %4ld (long %u, huge %u) entries of w %6ld : %s\n", name, (long) B.nnz, nl [0], nl [1], present,
        g_bad ? "FAILED" : "ok") ;
}

int main ()
{
    setvbuf (stdout, nullptr, _IOLBF, 0) ;
    // vectors of A: empty, shorter than a lane group, around a tile of the streamed SpMV (2048 entries), long
    // (> VEC_LONG = 2048: cut into segments), huge (> VEC_HUGE = 32768: spread over the grid), dense
    const int64_t n = 40000 ;
    std::vector<int64_t> lens = { 0, 1, 5, 31, 33, 0, 200, 2047, 2048, 2049, 7, 5000, 35000, 12, 40000, 64, 3, 900,
        0, 0, 17, 4100, 1, 33000, 260, 2, 2050, 100, 0, 9 } ;
    for (int rep = 0 ; rep < 3 ; rep++) for (int q = 0 ; q < 20 ; q++) lens.push_back ((q * 37 + rep * 11) % 150) ;
    std::mt19937 rng (77) ;
    {
        using S = Semiring<double, double, GB200_MIN, GB200_PLUS> ;
        Mat<double> A = make<double> (n, lens, rng) ;
        run_pull<S> ("MIN_PLUS_FP64 mxv streamed, dense u (SSSP)", A, true, 0, GB200_PLUS, 1) ;
        run_pull<S> ("MIN_PLUS_FP64 mxv streamed, sparse u", A, false, 0, GB200_PLUS, 2) ;
        run_pull<S> ("MIN_PLUS_FP64 mxv masked, sparse u", A, false, 1, GB200_PLUS, 3) ;
        run_push<S> ("MIN_PLUS_FP64 vxm push, no mask", A, 0.6, 0, GB200_PLUS, 4) ;
    }
    {
        using S = Semiring<double, double, GB200_PLUS, GB200_TIMES> ;
        Mat<double> A = make<double> (n, lens, rng) ;
        run_pull<S> ("PLUS_TIMES_FP64 mxv streamed, dense u", A, true, 0, GB200_TIMES, 5) ;
        run_pull<S> ("PLUS_TIMES_FP64 mxv complemented mask, dense u", A, true, 2, GB200_TIMES, 6) ;
        run_push<S> ("PLUS_TIMES_FP64 vxm push, mask", A, 0.5, 1, GB200_TIMES, 7) ;
    }
    {
        using S = Semiring<bool, bool, GB200_LOR, GB200_LAND> ;
        Mat<bool> A = make<bool> (n, lens, rng) ;
        run_pull<S> ("LOR_LAND_BOOL mxv complemented mask, sparse u (BFS pull)", A, false, 2, GB200_LAND, 8) ;
        run_pull<S> ("LOR_LAND_BOOL mxv streamed, sparse u", A, false, 0, GB200_LAND, 9) ;
        run_push<S> ("LOR_LAND_BOOL vxm push, complemented mask (BFS push)", A, 0.3, 2, GB200_LAND, 10) ;
    }
    printf ("emu_vec: %s\n", g_bad ? "FAILED" : "ok") ;
    return g_bad != 0 ;
}
