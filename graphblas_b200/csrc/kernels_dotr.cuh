// kernels_dotr.cuh -- the ROW-WALK masked dot kernel, C<M> = A'*B (included by kernels.cuh).
//
// Reference behaviour restated: Source/Template/GB_AxB_dot_mask.c:33-159 (one dot product per entry
// of M) with the inner loop of GB_AxB_dot_cij.c:47-256 (first product copied, later ones combined,
// entry emitted only if some index matched).
//
// Like dotg_kernel (kernels.cuh) the pairs arrive grouped by their OWNER (the longer vector), the
// owner goes into shared memory once per work item and the shorter list of every pair is walked
// against it.  What differs is who walks and what the owner is kept in:
//
//   * a WARP claims a batch of up to 32 tasks with one shared-memory atomic, lane l keeps the
//     descriptor of task l, and the warp walks the tasks one after another in ROWS of 32 consecutive
//     indices: one coalesced load, one branch-free probe per lane, one vote; four rows per iteration
//     with their loads in flight.
//   * regular owners (<= dotg_cap entries): the cuckoo tables of dotg_kernel.  A table that cannot be
//     built falls back to a binary search of the owner's list in global memory for that item.
//   * hub owners: a BITMAP over a part of the owner's index range in shared memory -- one shared
//     read and one bit test per probe, no build failures.  An owner whose range needs several parts is
//     served part by part: both lists are sorted, so every task keeps a 16-bit cursor (the row at
//     which it left the previous part) in shared memory, and an index is looked up in exactly one
//     part (an index outside [lo, lo+nbits) misses by the range test alone).
#pragma once

namespace gb200 {

constexpr int DOTR_THREADS = 512 ;              // cuckoo instantiation (64 KB table, 2-3 blocks per SM)
constexpr int DOTR_BM_THREADS = 1024 ;          // bitmap instantiation (one block per SM)
constexpr int DOTR_BM_BYTES = 184 * 1024 ;      // bitmap (pattern-only) or bitmap + rank (valued)
constexpr int DOTR_CUR_TASKS = DOTG_HUB_TASKS ; // cursors of one hub item: 2 bytes each
constexpr int DOTR_BM_SMEM = DOTR_BM_BYTES + 2 * DOTR_CUR_TASKS ;
constexpr int DOTR_HITS = 32 ;                  // valued operands: hits a warp collects before it loads their values
constexpr int DOTR_HIT_BYTES = DOTR_HITS * 8 ;  // per warp: (position in the owner, offset in the walk) pairs
constexpr int DOTR_MAXPARTS = 16 ;              // hub owners whose range needs more parts: dotg_kernel<HUB>
constexpr int DOTR_REBUILDS = 30 ;

// indices covered by one bitmap part
__host__ __device__ constexpr int64_t dotr_bm_bits (bool iso)
{
    // one word is kept back: bit `nbits` of a part must exist and be zero (dotr_probe)
    return iso ? (int64_t) DOTR_BM_BYTES * 8 - 32 : (int64_t) DOTR_BM_BYTES * 4 - 32 ;
}

enum { DOTR_CUCKOO = 0, DOTR_DENSE = 1, DOTR_BSEARCH = 2, DOTR_BITMAP = 3 } ;

template <class S> struct DotRCtx
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
    bool last ;                 // bitmap: this is the owner's last part
    int mode ;                  // cuckoo instantiation: DOTR_CUCKOO / DENSE / BSEARCH
    int NS, sh ;
    uint32_t c1, c2 ;
    uint32_t lo, nbits ;        // bitmap part: indices [lo, lo + nbits)
    int32_t hi ;                // lo + nbits
} ;

// The owner's table in shared memory, read through a 32-bit shared-memory address (ld.shared with a
// register base: no generic-to-shared address arithmetic in the probe).  word = 32-bit word index.
struct SmemTab
{
#ifdef GB200_HOST_EMULATION
    const uint32_t *base ;
    __device__ __forceinline__ explicit SmemTab (const void *p) : base ((const uint32_t *) p) { }
    __device__ __forceinline__ uint32_t ld (uint32_t word) const { return base [word] ; }
    __device__ __forceinline__ uint64_t ld64 (uint32_t dword) const { return ((const uint64_t *) base) [dword] ; }
#else
    uint32_t base ;
    __device__ __forceinline__ explicit SmemTab (const void *p) : base ((uint32_t) __cvta_generic_to_shared (p)) { }
    __device__ __forceinline__ uint32_t ld (uint32_t word) const
    {
        uint32_t v ;
        asm volatile ("ld.shared.u32 %0, [%1];" : "=r" (v) : "r" (base + (word << 2))) ;
        return v ;
    }
    __device__ __forceinline__ uint64_t ld64 (uint32_t dword) const
    {
        uint64_t v ;
        asm volatile ("ld.shared.u64 %0, [%1];" : "=l" (v) : "r" (base + (dword << 3))) ;
        return v ;
    }
#endif
} ;

// what a probe needs, in registers
struct DotRProbe
{
    uint32_t lo, nbits ;                // bitmap part: indices [lo, lo + nbits); bit nbits is zero
    int32_t hi ;                        // bitmap: lo + nbits, or INT32_MAX for the owner's last part
    uint32_t c1, c2 ; int sh ;          // cuckoo
    uint32_t NS ;                       // cuckoo: slots per table (the position tables start at 2 NS)
    const int32_t *Oi ; int olen ;      // BSEARCH
} ;

// one probe of index kq (NOKEY: a lane past the end of the list, never a hit); branch-free for the
// bitmap and the cuckoo tables
template <bool ISO, int MODE>
__device__ __forceinline__ uint32_t dotr_probe (const DotRProbe &q, const SmemTab &tab, const SmemTab &tab2,
    uint32_t kq, uint32_t &pos)
{
    constexpr uint32_t NOKEY = 0xFFFFFFFEu ;
    if constexpr (MODE == DOTR_BITMAP)
    {
        // an index outside [lo, lo + nbits) lands on bit `nbits`, which is kept zero (no range branch)
        uint32_t kk = kq - q.lo ;
        kk = (kk < q.nbits) ? kk : q.nbits ;
        const uint32_t w = kk >> 5 ;
        const uint32_t word = tab.ld (w) ;
        const uint32_t hit = (word >> (kk & 31)) & 1u ;
        if constexpr (!ISO)
            if (hit) pos = tab.ld ((DOTR_BM_BYTES / 8) + w) + __popc (word & ((1u << (kk & 31)) - 1u)) ;
        return hit ;
    }
    else if constexpr (MODE == DOTR_CUCKOO)
    {
        if constexpr (ISO)
        {
            const uint32_t e1 = tab.ld ((kq * q.c1) >> q.sh) ;
            const uint32_t e2 = tab2.ld ((kq * q.c2) >> q.sh) ;
            return ((e1 == kq) | (e2 == kq)) ? 1u : 0u ;
        }
        else
        {
            // valued operands: the same two key tables, and behind them two tables of positions that are
            // read on a hit only (64-bit (index, position) slots doubled the shared-memory traffic of
            // every probe, of which one in seven is a hit)
            const uint32_t h1 = (kq * q.c1) >> q.sh, h2 = (kq * q.c2) >> q.sh ;
            const uint32_t e1 = tab.ld (h1) ;
            const uint32_t e2 = tab2.ld (h2) ;
            const bool m1 = (e1 == kq), m2 = (e2 == kq) ;
            if (m1 | m2) pos = tab.ld (2u * q.NS + (m1 ? h1 : (q.NS + h2))) ;
            return (m1 | m2) ? 1u : 0u ;
        }
    }
    else if constexpr (MODE == DOTR_DENSE) { pos = kq ; return (kq != NOKEY) ? 1u : 0u ; }
    else
    {
        int l = 0, h = (kq != NOKEY) ? q.olen : 0 ;
        while (l < h)
        {
            const int mid = (l + h) >> 1 ;
            const uint32_t v = (uint32_t) __ldg (q.Oi + mid) ;
            if (v == kq) { pos = (uint32_t) mid ; return 1u ; }
            if (v < kq) l = mid + 1 ; else h = mid ;
        }
        return 0u ;
    }
}

// The warps of the block pull batches of tasks of the item until its counter runs out.  A task is
// walked in rows of 32 consecutive indices, DOTR_U_* rows per iteration with all their loads in flight.
// s_cur (BITMAP, several parts): per task of the item, the row at which the previous part left it.
constexpr int DOTR_U_TABLE = 4 ;      // rows per iteration against the cuckoo tables (tasks: ~5 rows)
constexpr int DOTR_U_BITMAP = 8 ;     // ... against a bitmap part (hub tasks: ~17 rows; one CTA per SM, so more
                                      // loads per warp have to be in flight)

template <class S, bool ISO, int MODE, class slot_t>
__device__ __forceinline__ void dotr_walk (const S &sr, const DotRCtx<S> &g, const void *table,
    int *s_next, uint16_t *s_cur, uint2 *hb, int nwarps, unsigned long long &nm)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    constexpr uint32_t NOKEY = 0xFFFFFFFEu ;
    constexpr unsigned FULL = 0xffffffffu ;
    constexpr bool BITMAP = (MODE == DOTR_BITMAP) ;
    constexpr int DOTR_U = BITMAP ? DOTR_U_BITMAP : DOTR_U_TABLE ;
    const int lane = threadIdx.x & 31 ;
    const SmemTab tab (table) ;
    DotRProbe q ;
    const SmemTab tab2 ((const char *) table + (size_t) g.NS * sizeof (uint32_t)) ;  // cuckoo: the second table
    q.lo = g.lo ; q.nbits = g.nbits ;
    q.hi = g.last ? INT32_MAX : g.hi ;
    q.c1 = g.c1 ; q.c2 = g.c2 ; q.sh = g.sh ; q.NS = (uint32_t) g.NS ; q.Oi = g.Oi ; q.olen = g.olen ;
    const bool cursors = BITMAP && g.multi ;
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
                int c = (nwarps > 0) ? rem / (2 * nwarps) : 32 ;        // nwarps == 0: the warp is alone
                c = (c < 2) ? 2 : ((c > 32) ? 32 : c) ;
                if (c > rem) c = rem ;
                const int seen = atomicCAS (s_next, cur, cur + c) ;
                if (seen == cur) { t0 = cur ; nb = c ; break ; }
                cur = seen ;
            }
        }
        t0 = __shfl_sync (FULL, t0, 0) ;
        nb = __shfl_sync (FULL, nb, 0) ;
        if (nb == 0) break ;
        // ---- lane l keeps task t0 + l: lc = its length | the row it resumes at << 16 -----------------
        uint32_t lc = 0 ;
        int32_t e = 0 ;
        long long w0 = 0 ;
        bool split = false ;
        if (lane < nb)
        {
            const DotTask d = g.tasks [t0 + lane] ;
            split = (d.len < 0) ;
            lc = (uint32_t) (split ? -d.len : d.len) ;          // <= DOTG_SEG
            e = d.e ; w0 = d.w0 ;
            if (cursors) lc |= ((uint32_t) s_cur [t0 + lane]) << 16 ;
        }
        uint32_t mycnt = 0 ;            // results of the task this lane keeps
        acc_t myacc = Mon::identity () ;
        uint32_t mycur = lc >> 16 ;
        for (int t = 0 ; t < nb ; t++)
        {
            const uint32_t tlc = __shfl_sync (FULL, lc, t) ;
            const int tl = (int) (tlc & 0xffffu), p0 = (int) (tlc >> 16) ;
            const long long tw = __shfl_sync (FULL, w0, t) ;
            uint32_t cnt = 0 ;
            acc_t acc = Mon::identity () ;
            bool found = false ;
            int stop = tl ;             // BITMAP: where the next part resumes
            int nbuf = 0 ;              // valued operands: hits noted in hb
            // the noted hits' values: lane l loads the pair of hit l and keeps its product
            auto flush = [&] (int n)
            {
                if constexpr (!ISO)
                {
                    __syncwarp () ;
                    if (lane < n)
                    {
                        const uint2 h = hb [lane] ;
                        const T ov = g.Ox [h.x], wv = g.Wx [tw + h.y] ;
                        const acc_t prod = g.orient ? sr.product (ov, wv) : sr.product (wv, ov) ;
                        acc = found ? Mon::combine (acc, prod) : prod ;
                        found = true ;
                    }
                    __syncwarp () ;
                }
            } ;
            if (p0 < tl)
            {
                const int32_t *__restrict__ rp = g.Wi + tw + p0 + lane ;    // this lane's index of row 0
                int rem = tl - p0 - lane ;                                  // > 32 u: an index in row u
                for (int p = p0 ; p < tl ; p += 32 * DOTR_U, rp += 32 * DOTR_U, rem -= 32 * DOTR_U)
                {
                    uint32_t k [DOTR_U] ;
                    #pragma unroll
                    for (int u = 0 ; u < DOTR_U ; u++)
                        k [u] = (rem > 32 * u) ? (uint32_t) __ldg (rp + 32 * u) : NOKEY ;
                    // rows 0 and 1 are probed whatever they hold (a row past the end is all NOKEY: no
                    // hit); every later pair of rows only if the task reaches it
                    #pragma unroll
                    for (int u = 0 ; u < DOTR_U ; u++)
                    {
                        if (u >= 2 && (u & 1) == 0 && p + 32 * u >= tl) break ;  // warp-uniform
                        uint32_t pos = 0 ;
                        const uint32_t hit = dotr_probe<ISO, MODE> (q, tab, tab2, k [u], pos) ;
                        cnt += hit ;
                        if constexpr (!ISO)
                        {
                            // valued operands: a hit is only NOTED here (where it is in the owner, where in
                            // the walk); the values of up to 32 hits are loaded together by flush below.
                            // Loading them row by row stalls the warp on two dependent loads per row with
                            // ~4 of 32 lanes at work (measured: 3x the pattern-only time).
                            const unsigned hm = __ballot_sync (FULL, hit) ;
                            if (hm)
                            {
                                const int c = __popc (hm) ;
                                if (nbuf + c > DOTR_HITS) { flush (nbuf) ; nbuf = 0 ; }
                                if (hit) hb [nbuf + __popc (hm & ((1u << lane) - 1u))] = make_uint2 (pos, (uint32_t) (p + 32 * u + lane)) ;
                                nbuf += c ;
                            }
                        }
                    }
                    if constexpr (BITMAP)
                    {
                        // The lists are sorted: an index at or above hi ends this part's stretch (such an
                        // index missed by the range test; hi = INT32_MAX in the owner's last part; NOKEY is
                        // negative).  The next part resumes at the first row that holds one.
                        int32_t mx = (int32_t) k [0] ;
                        #pragma unroll
                        for (int u = 1 ; u < DOTR_U ; u++) mx = ((int32_t) k [u] > mx) ? (int32_t) k [u] : mx ;
                        if (__any_sync (FULL, mx >= q.hi))
                        {
                            stop = p + 32 * (DOTR_U - 1) ;
                            #pragma unroll
                            for (int u = DOTR_U - 2 ; u >= 0 ; u--)
                                if (__any_sync (FULL, (int32_t) k [u] >= q.hi)) stop = p + 32 * u ;
                            break ;
                        }
                    }
                    if constexpr (!ISO)
                        if (Mon::has_terminal ())
                        {
                            if (nbuf > 0) { flush (nbuf) ; nbuf = 0 ; }
                            if (__any_sync (FULL, found && Mon::is_terminal (acc))) break ;   // stop = tl: decided
                        }
                }
            }
            // ---- the task's result goes to the lane that keeps it --------------------------------
            if constexpr (!ISO) { if (nbuf > 0) flush (nbuf) ; }
            const uint32_t tot = __reduce_add_sync (FULL, cnt) ;
            if constexpr (!ISO)
            {
                if (tot)
                {
                    // only lanes that hold a product take part (the reference copies the first product and
                    // combines the later ones, GB_AxB_dot_cij.c:29-45: a pair whose only products are NaN
                    // is NaN under MIN / MAX, where combining with the identity would give +-Inf)
                    unsigned fm = __ballot_sync (FULL, found) ;
                    #pragma unroll
                    for (int off = 16 ; off > 0 ; off >>= 1)
                    {
                        const acc_t other = __shfl_down_sync (FULL, acc, off) ;
                        const bool of = (lane + off < 32) && ((fm >> (lane + off)) & 1u) ;
                        if (of) acc = ((fm >> lane) & 1u) ? Mon::combine (acc, other) : other ;
                        fm |= (fm >> off) ;
                    }
                    acc = __shfl_sync (FULL, acc, 0) ;
                    // a dense owner: the reference starts from the identity there (dot_cij.c:110,125,141),
                    // so NaN products are dropped by fmin / fmax even when there is nothing else
                    if constexpr (MODE == DOTR_DENSE) acc = Mon::combine (Mon::identity (), acc) ;
                }
                if (lane == t) myacc = acc ;
            }
            if (lane == t) { mycnt = tot ; mycur = (uint32_t) stop ; }
        }
        // ---- every lane writes the result of its task ----------------------------------------------
        if (lane < nb)
        {
            if (cursors) s_cur [t0 + lane] = (uint16_t) mycur ;
            if (mycnt)
            {
                if constexpr (ISO) myacc = iso_fold<Mon> (g.ciso, mycnt) ;
                if (split || g.multi) Mon::atomic_combine (g.vals + e, myacc) ;
                else g.vals [e] = myacc ;
                g.flags [e] = 1 ;
                nm += mycnt ;
            }
        }
    }
}

// one pass of the cuckoo build (keys only: two tables of NS 32-bit slots); returns through *s_fail
__device__ __forceinline__ void dotr_cuckoo_pass (uint32_t *tab, const int32_t *__restrict__ Oi, int slen,
    int NS, int sh, uint32_t c1, uint32_t c2, int *s_fail)
{
    constexpr uint32_t EMPTY = ~0u ;
    for (int t = threadIdx.x ; t < 2 * NS ; t += blockDim.x) tab [t] = EMPTY ;
    if (threadIdx.x == 0) *s_fail = 0 ;
    __syncthreads () ;
    for (int q = threadIdx.x ; q < slen ; q += blockDim.x)
    {
        uint32_t cur = (uint32_t) __ldg (Oi + q) ;
        int which = 0, n = 0 ;
        #pragma unroll 1
        for ( ; n < DOTG_MAXIT ; n++)
        {
            const uint32_t loc = which ? (NS + ((cur * c2) >> sh)) : ((cur * c1) >> sh) ;
            cur = atomicExch (tab + loc, cur) ;
            if (cur == EMPTY) break ;
            which ^= 1 ;                        // the evicted entry moves to its other table
        }
        if (n == DOTG_MAXIT) *s_fail = 1 ;
    }
    __syncthreads () ;
}

// valued operands: once the keys have settled, entry q of the owner writes its position next to the slot
// its key ended up in (tables 3 and 4: tab [2 NS + slot]); `stride` threads starting at `first` share the work
__device__ __forceinline__ void dotr_cuckoo_positions (uint32_t *tab, const int32_t *__restrict__ Oi, int slen,
    int NS, int sh, uint32_t c1, uint32_t c2, int first, int stride)
{
    for (int q = first ; q < slen ; q += stride)
    {
        const uint32_t k = (uint32_t) __ldg (Oi + q) ;
        const uint32_t h1 = (k * c1) >> sh ;
        const uint32_t slot = (tab [h1] == k) ? h1 : (NS + ((k * c2) >> sh)) ;
        tab [2 * NS + slot] = (uint32_t) q ;
    }
}

template <class S, bool ISO, bool BITMAP>
__global__ void __launch_bounds__ (BITMAP ? DOTR_BM_THREADS : DOTR_THREADS, BITMAP ? 1 : (ISO ? 3 : 2))
dotr_kernel (DotGArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    using slot_t = uint32_t ;
    constexpr int NW = (BITMAP ? DOTR_BM_THREADS : DOTR_THREADS) / 32 ;
    extern __shared__ __align__ (16) unsigned char dotr_raw [] ;
    __shared__ int64_t s_ws [33] ;
    __shared__ int s_next, s_fail ;
    __shared__ unsigned long long s_item ;
    __shared__ int s_q0, s_q1 ;
    constexpr int CAP = dotg_cap (ISO) ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bx = (const T *) a.B.x ;
    const int lane = threadIdx.x & 31 ;
    const bool orient = (a.orient != 0) ;
    const DMat &O = orient ? a.A : a.B ;        // owner matrix (probed)
    const DMat &W = orient ? a.B : a.A ;        // walked matrix
    const T *__restrict__ Oxb = orient ? Ax : Bx ;
    const int64_t vlen = a.A.vlen ;
    uint16_t *s_cur = (uint16_t *) (dotr_raw + DOTR_BM_BYTES) ;         // BITMAP only
    // valued operands: the warp's hit buffer, behind the table (and the cursors)
    uint2 *hb = ISO ? nullptr : (uint2 *) (dotr_raw + (BITMAP ? DOTR_BM_SMEM : DOTG_SMEM) + (threadIdx.x >> 5) * DOTR_HIT_BYTES) ;
    DotRCtx<S> g ;
    g.Wi = W.i ; g.Wx = orient ? Bx : Ax ;
    g.vals = (acc_t *) a.vals ; g.flags = a.flags ; g.orient = orient ;
    g.ciso = Mon::identity () ;
    if (ISO) g.ciso = sr.product (Ax [0], Bx [0]) ;
    g.multi = false ; g.last = true ; g.lo = 0 ; g.nbits = 0 ; g.hi = INT32_MAX ;
    g.mode = DOTR_CUCKOO ; g.NS = 0 ; g.sh = 0 ; g.c1 = 0 ; g.c2 = 0 ;
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
            g.mode = dense ? DOTR_DENSE : DOTR_CUCKOO ;
            if (!dense && olen > CAP) g.mode = DOTR_BSEARCH ;       // never scheduled here; stay correct
            if (g.mode == DOTR_CUCKOO)
            {
                int lg = 5 ;                    // 2^lg slots per table: total load between 3/16 and 3/8
                while (3 * (1 << lg) < 4 * olen) lg++ ;
                const int NS = 1 << lg, sh = 32 - lg ;
                uint32_t c1 = 0x9E3779B1u, c2 = 0x85EBCA6Bu ;
                for (int attempt = 0 ; ; attempt++)
                {
                    dotr_cuckoo_pass ((uint32_t *) dotr_raw, g.Oi, olen, NS, sh, c1, c2, &s_fail) ;
                    const bool failed = (s_fail != 0) ;
                    __syncthreads () ;          // everyone has read s_fail before it is reset
                    if (!failed) break ;
                    if (attempt >= DOTR_REBUILDS) { g.mode = DOTR_BSEARCH ; break ; }
                    c1 = (c1 * 0x01000193u + 0xFE94F82Au) | 1u ;
                    c2 = (c2 * 0x01000193u + 0x4A8BE922u) | 1u ;
                }
                g.NS = NS ; g.sh = sh ; g.c1 = c1 ; g.c2 = c2 ;
                if constexpr (!ISO)
                    if (g.mode == DOTR_CUCKOO)
                        dotr_cuckoo_positions ((uint32_t *) dotr_raw, g.Oi, olen, NS, sh, c1, c2, threadIdx.x, blockDim.x) ;
            }
            if (threadIdx.x == 0) s_next = 0 ;
            __syncthreads () ;
            if (g.mode == DOTR_CUCKOO) dotr_walk<S, ISO, DOTR_CUCKOO, slot_t> (sr, g, dotr_raw, &s_next, nullptr, hb, NW, nm) ;
            else if (g.mode == DOTR_DENSE) dotr_walk<S, ISO, DOTR_DENSE, slot_t> (sr, g, dotr_raw, &s_next, nullptr, hb, NW, nm) ;
            else dotr_walk<S, ISO, DOTR_BSEARCH, slot_t> (sr, g, dotr_raw, &s_next, nullptr, hb, NW, nm) ;
        }
        else
        {
            // ---- hub owner: bitmap parts over [omin, omax] ------------------------------------------
            uint32_t *bm = (uint32_t *) dotr_raw ;
            uint32_t *rk = bm + (DOTR_BM_BYTES / 8) ;
            const int64_t BITS = a.bm_bits ;    // dotr_bm_bits (ISO) unless a test asks for small parts
            const int64_t omin = __ldg (O.i + o0), omax = __ldg (O.i + o1 - 1) ;
            const int64_t lo0 = omin & ~(int64_t) 31 ;
            const int nparts = (int) ((omax - lo0) / BITS) + 1 ;
            g.multi = (nparts > 1) ;
            if (g.multi) for (int t = threadIdx.x ; t < g.ntask ; t += blockDim.x) s_cur [t] = 0 ;
            for (int part = 0 ; part < nparts ; part++)
            {
                const int64_t lo = lo0 + (int64_t) part * BITS ;
                const int64_t hi = (lo + BITS < omax + 1) ? (lo + BITS) : (omax + 1) ;
                const int nwords = (int) ((hi - lo + 32) >> 5) ;    // with the spare zero bit
                __syncthreads () ;              // the previous part's walkers are done
                for (int t = threadIdx.x ; t < nwords ; t += blockDim.x) bm [t] = 0u ;
                if (threadIdx.x < 2)
                {
                    // the owner's entries of this part are a contiguous stretch [s_q0, s_q1) of its list
                    const int64_t bound = threadIdx.x ? hi : lo ;
                    int l = 0, h = olen ;
                    while (l < h)
                    {
                        const int mid = (l + h) >> 1 ;
                        if ((int64_t) __ldg (g.Oi + mid) < bound) l = mid + 1 ; else h = mid ;
                    }
                    if (threadIdx.x) s_q1 = l ; else s_q0 = l ;
                }
                __syncthreads () ;
                {
                    const int q0 = s_q0, q1 = s_q1 ;
                    const uint32_t ulo = (uint32_t) lo ;
                    int q = q0 + threadIdx.x ;
                    // four loads in flight per thread
                    for ( ; q + 3 * (int) blockDim.x < q1 ; q += 4 * blockDim.x)
                    {
                        uint32_t kk [4] ;
                        #pragma unroll
                        for (int u = 0 ; u < 4 ; u++) kk [u] = (uint32_t) __ldg (g.Oi + q + u * (int) blockDim.x) - ulo ;
                        #pragma unroll
                        for (int u = 0 ; u < 4 ; u++) atomicOr (bm + (kk [u] >> 5), 1u << (kk [u] & 31)) ;
                    }
                    for ( ; q < q1 ; q += blockDim.x)
                    {
                        const uint32_t kk = (uint32_t) __ldg (g.Oi + q) - ulo ;
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
                g.lo = (uint32_t) lo ; g.nbits = (uint32_t) (hi - lo) ; g.hi = (int32_t) hi ;
                g.last = (part == nparts - 1) ;
                dotr_walk<S, ISO, DOTR_BITMAP, slot_t> (sr, g, dotr_raw, &s_next, s_cur, hb, NW, nm) ;
            }
        }
    }
    for (int off = 16 ; off > 0 ; off >>= 1) nm += __shfl_down_sync (0xffffffffu, nm, off) ;
    if (lane == 0 && nm) atomicAdd (a.nmatch, nm) ;
}

// ---------------------------------------------------------------------------------------------
// TINY owners (DOTG_SMALL <= entries <= dotr_tiny_cap): a WARP, not a block, takes the work item.
// Such owners have a handful of tasks each (tri, RMAT scale 22: ~280 K of them per orientation, 6 to
// 60 tasks each), so a block per owner spends its time in barriers and in the latency of the item's
// set-up (measured: 3.2 warps stalled on the barrier per issued instruction in the block kernel).  Here
// every warp builds the cuckoo tables of its own owner in its own 1/16 of the block's shared memory
// (warp-synchronous, no block barrier) and walks the owner's tasks with dotr_walk.
// ---------------------------------------------------------------------------------------------
constexpr int DOTR_WARP_BYTES = DOTG_SMEM / (DOTR_THREADS / 32) ;       // table bytes per warp

__host__ __device__ constexpr int dotr_tiny_cap (bool iso) { return (DOTR_WARP_BYTES / (iso ? 4 : 8)) * 3 / 8 ; }

template <class S, bool ISO>
__global__ void __launch_bounds__ (DOTR_THREADS, ISO ? 3 : 2)
dotr_warp_kernel (DotGArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    using slot_t = uint32_t ;
    constexpr slot_t EMPTY = ~0u ;
    constexpr unsigned FULL = 0xffffffffu ;
    extern __shared__ __align__ (16) unsigned char dotr_raw [] ;
    __shared__ int s_wnext [DOTR_THREADS / 32] ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bx = (const T *) a.B.x ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5 ;
    const bool orient = (a.orient != 0) ;
    const DMat &O = orient ? a.A : a.B ;        // owner matrix (probed)
    const DMat &W = orient ? a.B : a.A ;        // walked matrix
    const T *__restrict__ Oxb = orient ? Ax : Bx ;
    slot_t *tab = (slot_t *) (dotr_raw + warp * DOTR_WARP_BYTES) ;
    uint2 *hb = ISO ? nullptr : (uint2 *) (dotr_raw + DOTG_SMEM + warp * DOTR_HIT_BYTES) ;
    DotRCtx<S> g ;
    g.Wi = W.i ; g.Wx = orient ? Bx : Ax ;
    g.vals = (acc_t *) a.vals ; g.flags = a.flags ; g.orient = orient ;
    g.ciso = Mon::identity () ;
    if (ISO) g.ciso = sr.product (Ax [0], Bx [0]) ;
    g.multi = false ; g.last = true ; g.lo = 0 ; g.nbits = 0 ; g.hi = INT32_MAX ;
    unsigned long long nm = 0 ;
    while (true)
    {
        unsigned long long it = 0 ;
        if (lane == 0) it = atomicAdd (a.next_item, 1ULL) ;
        it = __shfl_sync (FULL, it, 0) ;
        if ((int64_t) it >= a.nitems) break ;
        const DotItem item = a.items [it] ;
        int64_t ko = item.owner ;
        if (!orient) ko = dm_vecpos (a.B, dm_vecname (a.M, item.owner)) ;
        const int64_t o0 = __ldg (O.p + ko), o1 = __ldg (O.p + ko + 1) ;
        const int olen = (int) (o1 - o0) ;
        g.tasks = a.tasks + item.e0 ;
        g.ntask = (int) (item.e1 - item.e0) ;
        g.Ox = Oxb + o0 ; g.Oi = O.i + o0 ; g.olen = olen ;
        // ---- the warp's cuckoo tables: 2^lg slots each, total load between 3/16 and 3/8 ---------------
        int lg = 3 ;
        while (3 * (1 << lg) < 4 * olen) lg++ ;
        const int NS = 1 << lg, sh = 32 - lg ;
        uint32_t c1 = 0x9E3779B1u, c2 = 0x85EBCA6Bu ;
        g.mode = DOTR_CUCKOO ;
        if (olen > dotr_tiny_cap (ISO)) g.mode = DOTR_BSEARCH ;         // never scheduled here; stay correct
        for (int attempt = 0 ; g.mode == DOTR_CUCKOO ; attempt++)
        {
            __syncwarp () ;                     // the previous walk is over
            for (int t = lane ; t < 2 * NS ; t += 32) tab [t] = EMPTY ;
            __syncwarp () ;
            bool bad = false ;
            for (int q = lane ; q < olen ; q += 32)
            {
                uint32_t cur = (uint32_t) __ldg (g.Oi + q) ;
                int which = 0, n = 0 ;
                #pragma unroll 1
                for ( ; n < DOTG_MAXIT ; n++)
                {
                    const uint32_t loc = which ? (NS + ((cur * c2) >> sh)) : ((cur * c1) >> sh) ;
                    cur = atomicExch (tab + loc, cur) ;
                    if (cur == EMPTY) break ;
                    which ^= 1 ;                // the evicted entry moves to its other table
                }
                if (n == DOTG_MAXIT) bad = true ;
            }
            __syncwarp () ;
            if (!__any_sync (FULL, bad)) break ;
            if (attempt >= DOTR_REBUILDS) { g.mode = DOTR_BSEARCH ; break ; }
            c1 = (c1 * 0x01000193u + 0xFE94F82Au) | 1u ;
            c2 = (c2 * 0x01000193u + 0x4A8BE922u) | 1u ;
        }
        g.NS = NS ; g.sh = sh ; g.c1 = c1 ; g.c2 = c2 ;
        if constexpr (!ISO)
            if (g.mode == DOTR_CUCKOO) dotr_cuckoo_positions (tab, g.Oi, olen, NS, sh, c1, c2, lane, 32) ;
        if (lane == 0) s_wnext [warp] = 0 ;
        __syncwarp () ;
        if (g.mode == DOTR_CUCKOO) dotr_walk<S, ISO, DOTR_CUCKOO, slot_t> (sr, g, tab, s_wnext + warp, nullptr, hb, 0, nm) ;
        else dotr_walk<S, ISO, DOTR_BSEARCH, slot_t> (sr, g, tab, s_wnext + warp, nullptr, hb, 0, nm) ;
    }
    for (int off = 16 ; off > 0 ; off >>= 1) nm += __shfl_down_sync (FULL, nm, off) ;
    if (lane == 0 && nm) atomicAdd (a.nmatch, nm) ;
}

} // namespace gb200
