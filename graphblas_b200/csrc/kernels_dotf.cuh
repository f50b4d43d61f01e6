// kernels_dotf.cuh -- the FLAT masked dot kernel, C<M> = A'*B (included by kernels.cuh).
//
// Reference behaviour restated: Source/Template/GB_AxB_dot_mask.c:33-159 (one dot product per entry
// of M) with the inner loop of GB_AxB_dot_cij.c:47-256 (first product copied, later ones combined,
// entry emitted only if some index matched).
//
// Why it exists: dotg_kernel (kernels.cuh) gives a task (one mask entry, or one 1024-long piece of
// its walked list) to a WARP (regular owners) or to a LANE (hub owners).  After the trim the median
// walked list is far shorter than 4 x 32 indices, so most lanes of a warp-per-task walk idle, and a
// lane-per-task walk issues one uncoalesced 32-byte load per lane per 8 probes.  Both kernels were
// bound by the L1/shared-memory pipe at ~35 % useful lanes (profiles/r1).
//
// Here the walked lists of up to 32 consecutive tasks are laid end to end and the warp walks that
// FLAT index space 32 positions per step: every lane probes every step (but the last), the loads of
// a task are coalesced, and the per-task result needs no atomics and no shuffles for pattern-only
// operands -- the lane that owns task t counts, from the ballot of the hits, the bits of its own
// stretch of the window.
//
//   owner table, regular owners (<= dotg_cap entries): the cuckoo tables of dotg_kernel, same build.
//     A table that cannot be built in DOTG_REBUILDS attempts no longer traps: the item falls back to
//     a binary search of the owner's list in global memory (mode BSEARCH), so the context survives.
//   owner table, hub owners: a BITMAP over the owner's index range in shared memory (216 KB: 1.77 M
//     indices per part pattern-only, half of that with the rank array valued operands need for the
//     position of a hit).  One shared-memory read and one bit test per probe, no build failures; an
//     owner whose range needs several parts is served part by part, every task restricted to the
//     part's index range by two binary searches of its (sorted) walked list.
#pragma once

namespace gb200 {

constexpr int DOTF_THREADS = 512 ;              // cuckoo instantiation (64 KB table, 2-3 blocks per SM)
constexpr int DOTF_BM_THREADS = 1024 ;          // bitmap instantiation (one block per SM)
constexpr int DOTF_BM_SMEM = 216 * 1024 ;       // bitmap (+ rank) bytes
constexpr int DOTF_MAXPARTS = 8 ;               // hub owners whose range needs more parts: dotg_kernel<HUB>
constexpr int DOTF_U = 4 ;                      // windows of 32 positions with their loads in flight
constexpr int DOTF_REBUILDS = 30 ;

// indices covered by one bitmap part
__host__ __device__ constexpr int64_t dotf_bm_bits (bool iso)
{
    return iso ? (int64_t) DOTF_BM_SMEM * 8 : (int64_t) DOTF_BM_SMEM * 4 ;
}

enum { DOTF_CUCKOO = 0, DOTF_DENSE = 1, DOTF_BSEARCH = 2 } ;

template <class S> struct DotFCtx
{
    const DotTask *tasks ;      // of this item
    int ntask ;
    const int32_t *Wi ;         // walked matrix: indices, values
    const typename S::T *Wx ;
    const typename S::T *Ox ;   // owner values (of the whole owner vector)
    const int32_t *Oi ;         // owner indices (BSEARCH)
    int olen ;
    typename S::acc_t *vals ;
    uint8_t *flags ;
    typename S::acc_t ciso ;
    bool orient ;
    bool multi ;                // several parts: partial results meet in the accumulator
    // cuckoo
    int mode ;
    int NS, sh ;
    uint32_t c1, c2 ;
    // bitmap part: indices [lo, lo + nbits); restrict: tasks are cut to [vlo, vhi)
    uint32_t lo, nbits ;
    int64_t vlo, vhi ;
    bool cut_lo, cut_hi ;
} ;

// One warp walks batches of up to 32 tasks of the item until the item's task counter runs out.
// s_base: 32 words of shared memory of this warp.
template <class S, bool ISO, bool BITMAP, class slot_t>
__device__ __forceinline__ void dotf_run (const S &sr, const DotFCtx<S> &g, const void *table,
    int *s_next, long long *s_base, int nwarps, unsigned long long &nm)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    constexpr uint32_t NOKEY = 0xFFFFFFFEu ;
    const int lane = threadIdx.x & 31 ;
    const unsigned le = 0xffffffffu >> (31 - lane) ;            // lanes 0..lane
    const slot_t *__restrict__ tab = (const slot_t *) table ;
    const slot_t *__restrict__ tab2 = tab + g.NS ;
    const uint32_t *__restrict__ bm = (const uint32_t *) table ;
    const uint32_t *__restrict__ rk = bm + (DOTF_BM_SMEM / 8) ;  // valued bitmap parts only
    while (true)
    {
        // ---- claim a batch: 32 tasks while the item is long, fewer towards its end (the warps of a
        // block finish an item together) -------------------------------------------------------------
        int t0 = 0, nb = 0 ;
        if (lane == 0)
        {
            int cur = *((volatile int *) s_next) ;
            while (cur < g.ntask)
            {
                const int rem = g.ntask - cur ;
                int c = rem / (2 * nwarps) ;
                c = (c < 4) ? 4 : ((c > 32) ? 32 : c) ;
                if (c > rem) c = rem ;
                const int seen = atomicCAS (s_next, cur, cur + c) ;
                if (seen == cur) { t0 = cur ; nb = c ; break ; }
                cur = seen ;
            }
        }
        t0 = __shfl_sync (0xffffffffu, t0, 0) ;
        nb = __shfl_sync (0xffffffffu, nb, 0) ;
        if (nb == 0) break ;
        // ---- lane l owns task t0 + l -------------------------------------------------------------
        int len = 0 ;
        int32_t e = 0 ;
        long long w0 = 0 ;
        bool split = false ;
        if (lane < nb)
        {
            const DotTask d = g.tasks [t0 + lane] ;
            split = (d.len < 0) ;
            len = split ? -d.len : d.len ;
            e = d.e ; w0 = d.w0 ;
        }
        int a0 = 0, a1 = len ;
        if constexpr (BITMAP)
        {
            // several parts: the stretch of the walked list whose indices lie in this part
            if (g.cut_lo && len > 0)
            {
                int l = 0, h = len ;
                while (l < h)
                {
                    const int mid = (l + h) >> 1 ;
                    if ((int64_t) __ldg (g.Wi + w0 + mid) < g.vlo) l = mid + 1 ; else h = mid ;
                }
                a0 = l ;
            }
            if (g.cut_hi && len > 0)
            {
                int l = a0, h = len ;
                while (l < h)
                {
                    const int mid = (l + h) >> 1 ;
                    if ((int64_t) __ldg (g.Wi + w0 + mid) < g.vhi) l = mid + 1 ; else h = mid ;
                }
                a1 = l ;
            }
        }
        const int n = a1 - a0 ;
        int incl = n ;
        #pragma unroll
        for (int off = 1 ; off < 32 ; off <<= 1)
        {
            const int o = __shfl_up_sync (0xffffffffu, incl, off) ;
            if (lane >= off) incl += o ;
        }
        const int excl = incl - n ;
        const int total = __shfl_sync (0xffffffffu, incl, 31) ;
        const bool ne = (n > 0) ;
        const unsigned nemask = __ballot_sync (0xffffffffu, ne) ;
        // flat position f of the batch is index  base + f  of the walked matrix, base by task rank
        // (rank = position among the non-empty tasks of the batch)
        if (ne) s_base [__popc (nemask & (le >> 1))] = w0 + a0 - excl ;
        __syncwarp () ;
        uint32_t cnt = 0 ;
        acc_t acc = Mon::identity () ;
        bool found = false ;
        int before = 0 ;                        // non-empty tasks that start before the window
        for (int f0 = 0 ; f0 < total ; f0 += 32 * DOTF_U)
        {
            uint32_t k [DOTF_U] ;
            unsigned smask [DOTF_U] ;
            long long pw [DOTF_U] ;
            #pragma unroll
            for (int u = 0 ; u < DOTF_U ; u++)
            {
                const int fw = f0 + 32 * u ;
                const unsigned rel = (unsigned) (excl - fw) ;
                const unsigned bit = (ne && rel < 32u) ? (1u << rel) : 0u ;
                smask [u] = __reduce_or_sync (0xffffffffu, bit) ;       // task starts inside the window
                const int tr = before + __popc (smask [u] & le) - 1 ;
                before += __popc (smask [u]) ;
                const int f = fw + lane ;
                k [u] = NOKEY ; pw [u] = 0 ;
                if (f < total)
                {
                    pw [u] = s_base [tr] + f ;
                    k [u] = (uint32_t) __ldg (g.Wi + pw [u]) ;
                }
            }
            #pragma unroll
            for (int u = 0 ; u < DOTF_U ; u++)
            {
                const int fw = f0 + 32 * u ;
                if (u > 0 && fw >= total) continue ;            // warp-uniform
                const uint32_t kq = k [u] ;
                bool hit ;
                uint32_t pos = 0 ;
                if constexpr (BITMAP)
                {
                    const uint32_t kk = kq - g.lo ;
                    hit = false ;
                    if (kk < g.nbits)
                    {
                        const uint32_t word = bm [kk >> 5] ;
                        hit = (word >> (kk & 31)) & 1u ;
                        if constexpr (!ISO)
                            if (hit) pos = rk [kk >> 5] + __popc (word & ((1u << (kk & 31)) - 1u)) ;
                    }
                }
                else
                {
                    if (g.mode == DOTF_CUCKOO)
                        hit = dotg_probe<ISO, false, slot_t> (tab, tab2, kq, g.sh, g.c1, g.c2, pos) ;
                    else if (g.mode == DOTF_DENSE) { hit = (kq != NOKEY) ; pos = kq ; }
                    else
                    {
                        int l = 0, h = g.olen ;
                        hit = false ;
                        while (l < h && kq != NOKEY)
                        {
                            const int mid = (l + h) >> 1 ;
                            const uint32_t v = (uint32_t) __ldg (g.Oi + mid) ;
                            if (v == kq) { hit = true ; pos = (uint32_t) mid ; break ; }
                            if (v < kq) l = mid + 1 ; else h = mid ;
                        }
                    }
                }
                const unsigned hitmask = __ballot_sync (0xffffffffu, hit) ;
                // the stretch [wlo, whi) of this window that belongs to the task this lane owns
                const int wlo = ((excl > fw) ? excl : fw) - fw ;
                const int whi = ((incl < fw + 32) ? incl : (fw + 32)) - fw ;
                const bool in = (whi > wlo) ;                   // false for an empty task
                const unsigned m = in ? ((0xffffffffu >> (32 - whi)) & (0xffffffffu << wlo)) : 0u ;
                if constexpr (ISO) cnt += __popc (hitmask & m) ;
                else
                {
                    acc_t v = Mon::identity () ;
                    if (hit)
                    {
                        const T ov = g.Ox [pos], wv = g.Wx [pw [u]] ;
                        v = g.orient ? sr.product (ov, wv) : sr.product (wv, ov) ;
                    }
                    // inclusive scan over the lanes, segmented by task, of the products of the hit lanes;
                    // which lanes hold a product is known to every lane from the ballot
                    const unsigned sb = smask [u] & le ;
                    const int segstart = sb ? (31 - __clz (sb)) : 0 ;
                    #pragma unroll
                    for (int off = 1 ; off < 32 ; off <<= 1)
                    {
                        const acc_t v2 = __shfl_up_sync (0xffffffffu, v, off) ;
                        const int src = lane - off ;
                        if (src >= segstart)
                        {
                            const int rlo = (src - off + 1 > segstart) ? (src - off + 1) : segstart ;
                            const int mlo = (lane - off + 1 > segstart) ? (lane - off + 1) : segstart ;
                            const bool has2 = (hitmask & (0xffffffffu >> (31 - src)) & (0xffffffffu << rlo)) != 0 ;
                            const bool has1 = (hitmask & le & (0xffffffffu << mlo)) != 0 ;
                            if (has2) v = has1 ? Mon::combine (v2, v) : v2 ;
                        }
                    }
                    const acc_t vv = __shfl_sync (0xffffffffu, v, in ? (whi - 1) : lane) ;
                    const unsigned hm = hitmask & m ;
                    if (hm)
                    {
                        acc = found ? Mon::combine (acc, vv) : vv ;
                        found = true ;
                        cnt += __popc (hm) ;
                    }
                }
            }
        }
        __syncwarp () ;                         // s_base is rewritten by the next batch
        if (cnt)
        {
            if constexpr (ISO) acc = iso_fold<Mon> (g.ciso, cnt) ;
            if (split || g.multi) Mon::atomic_combine (g.vals + e, acc) ;
            else g.vals [e] = acc ;
            g.flags [e] = 1 ;
            nm += cnt ;
        }
    }
}

// one pass of the cuckoo build (the same as dotg_kernel's); returns through *s_fail
template <bool ISO, class slot_t>
__device__ __forceinline__ void dotf_cuckoo_pass (slot_t *tab, const int32_t *__restrict__ Oi, int slen,
    int NS, int sh, uint32_t c1, uint32_t c2, int *s_fail)
{
    constexpr slot_t EMPTY = (slot_t) ~(slot_t) 0 ;
    for (int t = threadIdx.x ; t < 2 * NS ; t += blockDim.x) tab [t] = EMPTY ;
    if (threadIdx.x == 0) *s_fail = 0 ;
    __syncthreads () ;
    for (int q = threadIdx.x ; q < slen ; q += blockDim.x)
    {
        slot_t cur ;
        if constexpr (ISO) cur = (uint32_t) __ldg (Oi + q) ;
        else cur = ((uint64_t) (uint32_t) q << 32) | (uint32_t) __ldg (Oi + q) ;
        int which = 0, n = 0 ;
        #pragma unroll 1
        for ( ; n < DOTG_MAXIT ; n++)
        {
            const uint32_t k = (uint32_t) cur ;
            const uint32_t loc = which ? (NS + ((k * c2) >> sh)) : ((k * c1) >> sh) ;
            if constexpr (ISO) cur = atomicExch (tab + loc, cur) ;
            else cur = atomicExch ((unsigned long long *) tab + loc, (unsigned long long) cur) ;
            if (cur == EMPTY) break ;
            which ^= 1 ;                        // the evicted entry moves to its other table
        }
        if (n == DOTG_MAXIT) *s_fail = 1 ;
    }
    __syncthreads () ;
}

template <class S, bool ISO, bool BITMAP>
__global__ void __launch_bounds__ (BITMAP ? DOTF_BM_THREADS : DOTF_THREADS, BITMAP ? 1 : (ISO ? 3 : 2))
dotf_kernel (DotGArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    using slot_t = typename std::conditional<ISO, uint32_t, uint64_t>::type ;
    constexpr int NW = (BITMAP ? DOTF_BM_THREADS : DOTF_THREADS) / 32 ;
    extern __shared__ __align__ (16) unsigned char dotf_raw [] ;
    __shared__ long long s_base [NW * 32] ;
    __shared__ int64_t s_ws [33] ;
    __shared__ int s_next, s_fail ;
    __shared__ unsigned long long s_item ;
    __shared__ int s_q0 ;
    constexpr int CAP = dotg_cap (ISO) ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bx = (const T *) a.B.x ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5 ;
    const bool orient = (a.orient != 0) ;
    const DMat &O = orient ? a.A : a.B ;        // owner matrix (probed)
    const DMat &W = orient ? a.B : a.A ;        // walked matrix
    const T *__restrict__ Oxb = orient ? Ax : Bx ;
    const int64_t vlen = a.A.vlen ;
    DotFCtx<S> g ;
    g.Wi = W.i ; g.Wx = orient ? Bx : Ax ;
    g.vals = (acc_t *) a.vals ; g.flags = a.flags ; g.orient = orient ;
    g.ciso = Mon::identity () ;
    if (ISO) g.ciso = sr.product (Ax [0], Bx [0]) ;
    g.multi = false ; g.cut_lo = false ; g.cut_hi = false ; g.vlo = 0 ; g.vhi = 0 ;
    g.lo = 0 ; g.nbits = 0 ; g.mode = DOTF_CUCKOO ; g.NS = 0 ; g.sh = 0 ; g.c1 = 0 ; g.c2 = 0 ;
    unsigned long long nm = 0 ;
    while (true)
    {
        __syncthreads () ;
        if (threadIdx.x == 0) s_item = atomicAdd (a.next_item, 1ULL) ;
        __syncthreads () ;
        const int64_t it = (int64_t) s_item ;
        if (it >= a.nitems) break ;
        const DotItem item = a.items [it] ;
        int64_t ko = item.owner ;
        if (!orient) ko = dm_vecpos (a.B, dm_vecname (a.M, item.owner)) ;
        const int64_t o0 = __ldg (O.p + ko), o1 = __ldg (O.p + ko + 1) ;
        const int olen = (int) (o1 - o0) ;
        g.tasks = a.tasks + item.e0 ;
        g.ntask = (int) (item.e1 - item.e0) ;
        g.Ox = Oxb + o0 ; g.Oi = O.i + o0 ; g.olen = olen ;
        if constexpr (!BITMAP)
        {
            // ---- regular owner: cuckoo tables (or nothing at all for a dense owner) -----------------
            const bool dense = ((int64_t) olen == vlen) ;
            g.mode = dense ? DOTF_DENSE : DOTF_CUCKOO ;
            if (!dense && olen > CAP) g.mode = DOTF_BSEARCH ;       // never scheduled here; stay correct
            if (g.mode == DOTF_CUCKOO)
            {
                int lg = 5 ;                    // 2^lg slots per table: total load between 3/16 and 3/8
                while (3 * (1 << lg) < 4 * olen) lg++ ;
                const int NS = 1 << lg, sh = 32 - lg ;
                uint32_t c1 = 0x9E3779B1u, c2 = 0x85EBCA6Bu ;
                for (int attempt = 0 ; ; attempt++)
                {
                    dotf_cuckoo_pass<ISO, slot_t> ((slot_t *) dotf_raw, g.Oi, olen, NS, sh, c1, c2, &s_fail) ;
                    const bool failed = (s_fail != 0) ;
                    __syncthreads () ;          // everyone has read s_fail before it is reset
                    if (!failed) break ;
                    if (attempt >= DOTF_REBUILDS) { g.mode = DOTF_BSEARCH ; break ; }
                    c1 = (c1 * 0x01000193u + 0xFE94F82Au) | 1u ;
                    c2 = (c2 * 0x01000193u + 0x4A8BE922u) | 1u ;
                }
                g.NS = NS ; g.sh = sh ; g.c1 = c1 ; g.c2 = c2 ;
            }
            if (threadIdx.x == 0) s_next = 0 ;
            __syncthreads () ;
            dotf_run<S, ISO, false, slot_t> (sr, g, dotf_raw, &s_next, s_base + warp * 32, NW, nm) ;
        }
        else
        {
            // ---- hub owner: bitmap parts over [omin, omax] ------------------------------------------
            uint32_t *bm = (uint32_t *) dotf_raw ;
            uint32_t *rk = bm + (DOTF_BM_SMEM / 8) ;
            const int64_t BITS = a.bm_bits ;    // dotf_bm_bits (ISO) unless a test asks for small parts
            const int64_t omin = __ldg (O.i + o0), omax = __ldg (O.i + o1 - 1) ;
            const int64_t lo0 = omin & ~(int64_t) 31 ;
            const int nparts = (int) ((omax - lo0) / BITS) + 1 ;
            g.multi = (nparts > 1) ;
            for (int part = 0 ; part < nparts ; part++)
            {
                const int64_t lo = lo0 + (int64_t) part * BITS ;
                const int64_t hi = (lo + BITS < omax + 1) ? (lo + BITS) : (omax + 1) ;
                const int nwords = (int) ((hi - lo + 31) >> 5) ;
                __syncthreads () ;              // the previous part's walkers are done
                for (int t = threadIdx.x ; t < nwords ; t += blockDim.x) bm [t] = 0u ;
                if (!ISO && threadIdx.x == 0)
                {
                    // owner entries before this part: the position of a hit counts from the owner's start
                    int l = 0, h = olen ;
                    while (l < h)
                    {
                        const int mid = (l + h) >> 1 ;
                        if ((int64_t) __ldg (g.Oi + mid) < lo) l = mid + 1 ; else h = mid ;
                    }
                    s_q0 = l ;
                }
                __syncthreads () ;
                for (int q = threadIdx.x ; q < olen ; q += blockDim.x)
                {
                    const int64_t k = __ldg (g.Oi + q) ;
                    if (k >= lo && k < hi)
                    {
                        const uint32_t kk = (uint32_t) (k - lo) ;
                        atomicOr (bm + (kk >> 5), 1u << (kk & 31)) ;
                    }
                }
                __syncthreads () ;
                if constexpr (!ISO)
                {
                    // rk [w] = owner entries before word w
                    const int per = (nwords + (int) blockDim.x - 1) / (int) blockDim.x ;
                    const int wa = threadIdx.x * per ;
                    const int wb = (wa + per < nwords) ? (wa + per) : nwords ;
                    int64_t mine = 0 ;
                    for (int t = wa ; t < wb ; t++) mine += __popc (bm [t]) ;
                    int64_t tot ;
                    int64_t run = s_q0 + block_excl_scan_i64 (mine, s_ws, tot) ;
                    for (int t = wa ; t < wb ; t++) { rk [t] = (uint32_t) run ; run += __popc (bm [t]) ; }
                }
                if (threadIdx.x == 0) s_next = 0 ;
                __syncthreads () ;
                g.lo = (uint32_t) lo ; g.nbits = (uint32_t) (hi - lo) ;
                g.vlo = lo ; g.vhi = hi ;
                g.cut_lo = (part > 0) ; g.cut_hi = (part < nparts - 1) ;
                dotf_run<S, ISO, true, slot_t> (sr, g, dotf_raw, &s_next, s_base + warp * 32, NW, nm) ;
            }
        }
    }
    for (int off = 16 ; off > 0 ; off >>= 1) nm += __shfl_down_sync (0xffffffffu, nm, off) ;
    if (lane == 0 && nm) atomicAdd (a.nmatch, nm) ;
}

} // namespace gb200
