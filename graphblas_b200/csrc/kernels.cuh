// kernels.cuh -- the semiring-templated kernels: saxpy numeric phase (C=A*B, C<M>=A*B) and the
// dot-product family (C<M>=A'*B, C<!M>=A'*B, C=A'*B).
//
// Reference behaviour restated (not translated):
//   saxpy numeric  : Source/Template/GB_AxB_Gustavson_nomask.c:66-159 (Work[i] (+)= A(i,k)(x)B(k,j))
//                    Source/Template/GB_AxB_Gustavson_mask.c:93-285   (only where M(i,j) is true)
//   dot            : Source/Template/GB_AxB_dot_cij.c:47-256 with the mask / complemented-mask /
//                    no-mask drivers dot_mask.c:33-159, dot_compmask.c:20-126, dot_nomask.c:20-83
// GPU design: the pattern of every output vector is known before the numeric phase (from the
// symbolic phase, or it is the mask's pattern), so an output entry is a *slot* found by binary
// search in a sorted index list (light vectors) or by bitmap rank (heavy vectors); products are
// combined into slots with the monoid's atomic.  No sort is needed and the kernels are identical
// for the masked and unmasked cases.
#pragma once
#include "common.cuh"
#include "semiring.cuh"
#include "kernels_vec.cuh"

namespace gb200 {

struct SaxpyArgs
{
    DMat A, B ;
    const int32_t *cols ;       // stored-vector positions kk of B handled by this launch
    int64_t ncols ;
    const int64_t *lp ;         // slot list: vector pointers ...
    const int32_t *li ;         // ... and sorted indices (pattern of C(:,j), or of M(:,j))
    const int64_t *lpos ;       // kk -> position in lp (nullptr: identity); -1: no such vector
    void *acc ;                 // accumulators, one per slot, pre-set to the monoid identity
    uint8_t *flags ;            // masked: set to 1 when a slot receives a product (else nullptr)
    int masked ;                // 1: an index that is not in the list is skipped
    // heavy vectors
    const HeavyItem *items ;
    int64_t nitems ;
    const uint32_t *bitmap ;    // nws bitmaps of nwords words
    const int32_t *rank ;       // nws arrays of nwords prefix popcounts
    int64_t nwords ;
    int mult_op ; int flip ;
} ;

// ---------------------------------------------------------------------------------------------
// saxpy numeric, light vectors: one thread block (32..512 threads) per vector of B.  Warps take
// entries B(k,j) round-robin, lanes stride over A(:,k).
// ---------------------------------------------------------------------------------------------
template <class S>
__global__ void saxpy_light_kernel (SaxpyArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bx = (const T *) a.B.x ;
    acc_t *__restrict__ acc = (acc_t *) a.acc ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5 ;
    for (int64_t c = blockIdx.x ; c < a.ncols ; c += gridDim.x)
    {
        const int64_t kk = a.cols [c] ;
        const int64_t lv = a.lpos ? a.lpos [kk] : kk ;
        if (lv < 0) continue ;
        const int64_t l0 = a.lp [lv], l1 = a.lp [lv+1] ;
        if (l1 <= l0) continue ;
        const int64_t pb0 = a.B.p [kk], pb1 = a.B.p [kk+1] ;
        for (int64_t pb = pb0 + warp ; pb < pb1 ; pb += nwarps)
        {
            const int64_t k = a.B.i [pb] ;
            int64_t pa, pe ;
            if (!dm_lookup (a.A, k, pa, pe)) continue ;
            const T bkj = Bx [pb] ;
            for (int64_t p = pa + lane ; p < pe ; p += 32)
            {
                const int32_t i = __ldg (a.A.i + p) ;
                const int64_t slot = bsearch_i32 (a.li, l0, l1, i) ;
                if (slot < 0) continue ;            // only possible when masked
                Mon::atomic_combine (acc + slot, sr.product (Ax [p], bkj)) ;
                if (a.flags) a.flags [slot] = 1 ;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// saxpy numeric, heavy vectors: many blocks per vector (one per chunk of B entries); the slot is
// base + rank of the row in the vector's bitmap.
// ---------------------------------------------------------------------------------------------
template <class S>
__global__ void saxpy_heavy_kernel (SaxpyArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bx = (const T *) a.B.x ;
    acc_t *__restrict__ acc = (acc_t *) a.acc ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5 ;
    for (int64_t it = blockIdx.x ; it < a.nitems ; it += gridDim.x)
    {
        const HeavyItem item = a.items [it] ;
        const int64_t lv = a.lpos ? a.lpos [item.kk] : item.kk ;
        if (lv < 0) continue ;
        const int64_t base = a.lp [lv] ;
        const uint32_t *__restrict__ bm = a.bitmap + (int64_t) item.w * a.nwords ;
        const int32_t  *__restrict__ rk = a.rank   + (int64_t) item.w * a.nwords ;
        for (int64_t pb = item.pb0 + warp ; pb < item.pb1 ; pb += nwarps)
        {
            const int64_t k = a.B.i [pb] ;
            int64_t pa, pe ;
            if (!dm_lookup (a.A, k, pa, pe)) continue ;
            const T bkj = Bx [pb] ;
            for (int64_t p = pa + lane ; p < pe ; p += 32)
            {
                const uint32_t i = (uint32_t) __ldg (a.A.i + p) ;
                const uint32_t word = __ldg (bm + (i >> 5)) ;
                const uint32_t bit = 1u << (i & 31) ;
                if (!(word & bit)) continue ;       // only possible when masked
                const int64_t slot = base + __ldg (rk + (i >> 5)) + __popc (word & (bit - 1)) ;
                Mon::atomic_combine (acc + slot, sr.product (Ax [p], bkj)) ;
                if (a.flags) a.flags [slot] = 1 ;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// dot products.  A group of G lanes (G = 1,2,4,..,32) computes one C(i,j) = A(:,i)' (x) B(:,j).
// ---------------------------------------------------------------------------------------------
enum { DOT_MASK = 0, DOT_COMP = 1, DOT_NONE = 2 } ;

struct DotArgs
{
    DMat A, B, M ;              // M: structural (false-valued entries already removed)
    int mode ;                  // DOT_MASK: pairs are the entries of M
                                // DOT_COMP / DOT_NONE: pairs are (ia, jb), ia < A.nvec, jb in [jb0,jb1)
    const int32_t *mvec ;       // DOT_MASK: stored-vector position in M of every entry of M
    const int64_t *mposB ;      // DOT_COMP: jb -> position of vector j in M.p, or -1
    int64_t jb0, jb1 ;
    int64_t npairs ;
    void *vals ;                // acc_t per pair
    uint8_t *flags ;            // 1 iff the pair has at least one matched index
    int G ;                     // lanes per pair
    unsigned long long *nmatch ; // total number of matched index pairs (the dot method's flops)
    int mult_op ; int flip ;
} ;

template <class S>
__global__ void dot_kernel (DotArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bx = (const T *) a.B.x ;
    acc_t *__restrict__ vals = (acc_t *) a.vals ;
    const int G = a.G ;
    const int gl = threadIdx.x & (G - 1) ;                      // lane within the group
    const int64_t groups_per_block = blockDim.x / G ;
    const int64_t gid0 = (int64_t) blockIdx.x * groups_per_block + threadIdx.x / G ;
    const int64_t gstride = (int64_t) gridDim.x * groups_per_block ;
    // every lane of a warp runs the same number of outer iterations so the shuffles are safe
    const int64_t niter = (a.npairs + gstride - 1) / gstride ;
    unsigned long long nm = 0 ;
    for (int64_t itn = 0 ; itn < niter ; itn++)
    {
        const int64_t e = gid0 + itn * gstride ;
        bool live = (e < a.npairs) ;
        int64_t pa = 0, pe = 0, pb = 0, pbe = 0 ;
        if (live)
        {
            int64_t i, jb ;
            if (a.mode == DOT_MASK)
            {
                i = a.M.i [e] ;
                const int64_t j = dm_vecname (a.M, a.mvec [e]) ;
                live = dm_lookup (a.B, j, pb, pbe) && dm_lookup (a.A, i, pa, pe) ;
            }
            else
            {
                const int64_t ia = e % a.A.nvec ;
                jb = a.jb0 + e / a.A.nvec ;
                pa = a.A.p [ia] ; pe = a.A.p [ia+1] ;
                pb = a.B.p [jb] ; pbe = a.B.p [jb+1] ;
                live = (pe > pa) && (pbe > pb) ;
                if (live && a.mode == DOT_COMP)
                {
                    const int64_t mv = a.mposB [jb] ;
                    if (mv >= 0)
                    {
                        i = dm_vecname (a.A, ia) ;
                        if (bsearch_i32 (a.M.i, a.M.p [mv], a.M.p [mv+1], (int32_t) i) >= 0)
                            live = false ;          // M(i,j) is true: C(i,j) is not computed
                    }
                }
            }
        }
        acc_t cij = Mon::identity () ;
        bool found = false ;
        if (live)
        {
            const int64_t ainz = pe - pa, bjnz = pbe - pb ;
            if (bjnz == a.B.vlen)
            {
                // B(:,j) is dense: every entry of A(:,i) matches (dot_cij.c:117-131)
                for (int64_t p = pa + gl ; p < pe ; p += G)
                {
                    const int64_t k = a.A.i [p] ;
                    cij = Mon::combine (cij, sr.product (Ax [p], Bx [pb + k])) ;
                    found = true ; nm++ ;
                    if (Mon::has_terminal () && Mon::is_terminal (cij)) break ;
                }
            }
            else if (ainz == a.A.vlen)
            {
                // A(:,i) is dense (dot_cij.c:133-148)
                for (int64_t p = pb + gl ; p < pbe ; p += G)
                {
                    const int64_t k = a.B.i [p] ;
                    cij = Mon::combine (cij, sr.product (Ax [pa + k], Bx [p])) ;
                    found = true ; nm++ ;
                    if (Mon::has_terminal () && Mon::is_terminal (cij)) break ;
                }
            }
            else if (ainz <= bjnz)
            {
                // walk the shorter list, binary-search the longer with a moving left bound
                int64_t lo = pb ;
                for (int64_t p = pa + gl ; p < pe ; p += G)
                {
                    const int32_t k = __ldg (a.A.i + p) ;
                    int64_t l = lo, h = pbe ;
                    while (l < h)
                    {
                        const int64_t mid = (l + h) >> 1 ;
                        if (__ldg (a.B.i + mid) < k) l = mid + 1 ; else h = mid ;
                    }
                    lo = l ;
                    if (l >= pbe) break ;
                    if (__ldg (a.B.i + l) == k)
                    {
                        cij = Mon::combine (cij, sr.product (Ax [p], Bx [l])) ;
                        found = true ; nm++ ;
                        if (Mon::has_terminal () && Mon::is_terminal (cij)) break ;
                    }
                }
            }
            else
            {
                int64_t lo = pa ;
                for (int64_t p = pb + gl ; p < pbe ; p += G)
                {
                    const int32_t k = __ldg (a.B.i + p) ;
                    int64_t l = lo, h = pe ;
                    while (l < h)
                    {
                        const int64_t mid = (l + h) >> 1 ;
                        if (__ldg (a.A.i + mid) < k) l = mid + 1 ; else h = mid ;
                    }
                    lo = l ;
                    if (l >= pe) break ;
                    if (__ldg (a.A.i + l) == k)
                    {
                        cij = Mon::combine (cij, sr.product (Ax [l], Bx [p])) ;
                        found = true ; nm++ ;
                        if (Mon::has_terminal () && Mon::is_terminal (cij)) break ;
                    }
                }
            }
        }
        // combine the G partial results (a fixed tree: deterministic for floating point)
        unsigned fm = __ballot_sync (0xffffffffu, found) ;
        for (int off = G >> 1 ; off > 0 ; off >>= 1)
        {
            acc_t other = __shfl_down_sync (0xffffffffu, cij, off, G) ;
            // a lane without any match holds the identity; identity (+) t == t for every monoid
            // (bit-for-bit except +0.0 + -0.0), so it can be combined unconditionally
            if (gl + off < G) cij = Mon::combine (cij, other) ;
        }
        if (e < a.npairs && gl == 0)
        {
            const int wl = threadIdx.x & 31 ;
            const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(G - 1))) ;
            const bool any = (fm & gmask) != 0 ;
            a.flags [e] = any ? 1 : 0 ;
            if (any) vals [e] = cij ;
        }
    }
    for (int off = 16 ; off > 0 ; off >>= 1) nm += __shfl_down_sync (0xffffffffu, nm, off) ;
    if ((threadIdx.x & 31) == 0 && nm) atomicAdd (a.nmatch, nm) ;
}

// ---------------------------------------------------------------------------------------------
// masked dot products, C<M> = A'*B, grouped by the LONGER vector of each pair ("owner").
//
// For a mask entry (i,j) the shorter of A(:,i), B(:,j) is walked and the longer one is probed.  On a
// power-law graph the walked lengths sum to ~10x the number of matches, so the probes must not go
// to DRAM: pairs are grouped by their owner vector, a thread block loads the owner into a
// shared-memory hash table once and then serves every task of its work item from it (owners longer
// than DOTG_CAP are probed through their global hash index, which stays L2-resident because all
// work items of one owner run back to back).  orient 0: owner = B(:,j); orient 1: owner = A(:,i)
// (the mask entries regrouped by i).  A TASK is one pair, or one DOTG_SEG-long segment of a pair
// whose walked list is longer than that (segments are combined with the monoid's atomic); the warps
// of a block pull tasks from a shared counter, so one long pair cannot stall a block.
// ---------------------------------------------------------------------------------------------
constexpr int DOTG_SLOTS = 8192 ;           // shared-memory hash slots per block (64 KB)
constexpr int DOTG_CAP = 4096 ;             // owners up to this long use the shared-memory table
constexpr int DOTG_SEG = 1024 ;             // longest walk of one task
constexpr int DOTG_THREADS = 512 ;

struct DotItem { int32_t owner ; int32_t pad ; int64_t e0, e1 ; } ;        // owner, task range
struct DotTask { int32_t e ; int32_t len ; int64_t w0 ; } ;                // len < 0: segment of a split pair

struct DotGArgs
{
    DMat A, B, M ;
    const DotTask *tasks ;
    const DotItem *items ;
    int64_t nitems ;
    int orient ;
    void *vals ;                // pre-set to the monoid identity
    uint8_t *flags ;            // pre-zeroed
    unsigned long long *nmatch ;
    unsigned long long *next_item ;     // dynamic work-item counter (zeroed before launch)
    int use_bloom ;
    int mult_op ; int flip ;
} ;

// stored-vector position of vector `name`, or -1
__device__ __forceinline__ int64_t dm_vecpos (const DMat &A, int64_t name)
{
    if (!A.hyper) return name ;
    int64_t lo = 0, hi = A.nvec - 1 ;
    while (lo <= hi)
    {
        const int64_t mid = (lo + hi) >> 1, hv = __ldg (A.h + mid) ;
        if (hv == name) return mid ;
        if (hv < name) lo = mid + 1 ; else hi = mid - 1 ;
    }
    return -1 ;
}

// true: walk A(:,i) and probe B(:,j) (owner B); false: walk B(:,j) and probe A(:,i) (owner A)
__device__ __forceinline__ bool dot_walkA (int64_t ainz, int64_t bjnz, int64_t vlen)
{
    return (bjnz == vlen) || (ainz != vlen && ainz <= bjnz) ;
}

template <class S>
__global__ void __launch_bounds__ (DOTG_THREADS, 3)
dotg_kernel (DotGArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    extern __shared__ int32_t dotg_sm [] ;
    int32_t *tkeys = dotg_sm ;
    int32_t *tpos = dotg_sm + DOTG_SLOTS ;
    __shared__ int s_next ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bx = (const T *) a.B.x ;
    acc_t *__restrict__ vals = (acc_t *) a.vals ;
    const int lane = threadIdx.x & 31 ;
    const DMat &O = a.orient ? a.A : a.B ;      // owner matrix (probed)
    const DMat &W = a.orient ? a.B : a.A ;      // walked matrix
    const int32_t *__restrict__ Wi = W.i ;
    const int64_t vlen = a.A.vlen ;
    unsigned long long nm = 0 ;
    __shared__ unsigned long long s_item ;
    while (true)
    {
        // work items vary in cost by orders of magnitude: blocks pull them from a global counter
        __syncthreads () ;
        if (threadIdx.x == 0) s_item = atomicAdd (a.next_item, 1ULL) ;
        __syncthreads () ;
        const int64_t it = (int64_t) s_item ;
        if (it >= a.nitems) break ;
        const DotItem item = a.items [it] ;
        // ---- the owner vector ---------------------------------------------------------------
        int64_t ko = item.owner ;
        if (!a.orient) ko = dm_vecpos (a.B, dm_vecname (a.M, item.owner)) ;
        const int64_t o0 = __ldg (O.p + ko), o1 = __ldg (O.p + ko + 1) ;
        const int64_t olen = o1 - o0 ;
        const int mode = (olen == vlen) ? 0 : ((olen <= DOTG_CAP) ? 1 : 2) ;
        const int64_t ohinfo = (mode == 2) ? __ldg (O.hinfo + ko) : -1 ;
        // mode 1: table of 4x the owner's length (load <= 0.25 keeps probe sequences short and the
        // clearing cost proportional to the owner); mode 2: all 64 KB are a Bloom filter in front of
        // the owner's global hash index, so that most misses never leave the SM
        int lg = 6 ;
        while ((1 << lg) < 4 * olen && lg < 13) lg++ ;
        const uint32_t tmask = (1u << lg) - 1u ;
        uint32_t *bloom = (uint32_t *) dotg_sm ;
        __syncthreads () ;                          // previous item's probes and counter are done
        if (threadIdx.x == 0) s_next = 0 ;
        if (mode == 1)
        {
            for (int t = threadIdx.x ; t <= (int) tmask ; t += blockDim.x) tkeys [t] = -1 ;
            __syncthreads () ;
            for (int64_t q = o0 + threadIdx.x ; q < o1 ; q += blockDim.x)
            {
                const int32_t key = __ldg (O.i + q) ;
                uint32_t h = hash32 ((uint32_t) key) >> (32 - lg) ;
                while (atomicCAS (tkeys + h, -1, key) != -1) h = (h + 1) & tmask ;
                tpos [h] = (int32_t) (q - o0) ;
            }
        }
        else if (mode == 2 && a.use_bloom)
        {
            for (int t = threadIdx.x ; t < 2 * DOTG_SLOTS ; t += blockDim.x) bloom [t] = 0u ;
            __syncthreads () ;
            for (int64_t q = o0 + threadIdx.x ; q < o1 ; q += blockDim.x)
            {
                const uint32_t b = hash32b ((uint32_t) __ldg (O.i + q)) >> 13 ;    // 19 bits
                atomicOr (bloom + (b >> 5), 1u << (b & 31)) ;
            }
        }
        __syncthreads () ;
        // ---- tasks of this item: warps pull them from a shared counter -----------------------
        const int ntask = (int) (item.e1 - item.e0) ;
        while (true)
        {
            int tk = 0 ;
            if (lane == 0) tk = atomicAdd (&s_next, 1) ;
            tk = __shfl_sync (0xffffffffu, tk, 0) ;
            if (tk >= ntask) break ;
            const DotTask task = a.tasks [item.e0 + tk] ;
            const bool split = (task.len < 0) ;
            const int64_t w0 = task.w0, w1 = task.w0 + (split ? -task.len : task.len) ;
            acc_t cij = Mon::identity () ;
            bool found = false ;
            for (int64_t p = w0 + lane ; p < w1 ; p += 32)
            {
                const int32_t k = __ldg (Wi + p) ;
                int64_t pos = -1 ;
                if (mode == 0) pos = o0 + k ;
                else if (mode == 1)
                {
                    uint32_t h = hash32 ((uint32_t) k) >> (32 - lg) ;
                    while (true)
                    {
                        const int32_t kk = tkeys [h] ;
                        if (kk == k) { pos = o0 + tpos [h] ; break ; }
                        if (kk < 0) break ;
                        h = (h + 1) & tmask ;
                    }
                }
                else
                {
                    const uint32_t b = hash32b ((uint32_t) k) >> 13 ;
                    if (!a.use_bloom || ((bloom [b >> 5] >> (b & 31)) & 1u)) pos = vechash_probe (O, ohinfo, o0, k) ;
                }
                if (pos >= 0)
                {
                    const acc_t prod = a.orient ? sr.product (Ax [pos], Bx [p]) : sr.product (Ax [p], Bx [pos]) ;
                    cij = found ? Mon::combine (cij, prod) : prod ;
                    found = true ; nm++ ;
                    if (Mon::has_terminal () && Mon::is_terminal (cij)) break ;
                }
            }
            const unsigned fm = __ballot_sync (0xffffffffu, found) ;
            if (fm == 0) continue ;                 // flags [e] stays 0
            if (!found) cij = Mon::identity () ;
            for (int off = 16 ; off > 0 ; off >>= 1)
            {
                const acc_t other = __shfl_down_sync (0xffffffffu, cij, off) ;
                cij = Mon::combine (cij, other) ;
            }
            if (lane == 0)
            {
                if (split) Mon::atomic_combine (vals + task.e, cij) ;
                else vals [task.e] = cij ;
                a.flags [task.e] = 1 ;
            }
        }
    }
    for (int off = 16 ; off > 0 ; off >>= 1) nm += __shfl_down_sync (0xffffffffu, nm, off) ;
    if (lane == 0 && nm) atomicAdd (a.nmatch, nm) ;
}

// ---------------------------------------------------------------------------------------------
// launchers, one set per (xy type); defined in inst_*.cu through GB200_INSTANTIATE_TYPE
// ---------------------------------------------------------------------------------------------
enum { FAM_SAXPY_LIGHT = 0, FAM_SAXPY_HEAVY = 1, FAM_DOT = 2, FAM_DOTG = 3, FAM_DOTV = 4,
    FAM_DOTV_LONG = 5, FAM_SAXPYV = 6, FAM_SAXPYV_LONG = 7, FAM_SPMV = 8, FAM_SPMV_PRES = 9 } ;

struct LaunchCfg { int grid ; int block ; cudaStream_t stream ; } ;

// returns false if (z_code, add) is not a built-in combination for this xy type
typedef bool (*launch_fn) (int family, int z_code, int add_opcode, const void *args,
    LaunchCfg cfg) ;

template <class S>
inline void launch_family (int family, const void *args, LaunchCfg cfg)
{
    if (family == FAM_SAXPY_LIGHT)
        saxpy_light_kernel<S> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const SaxpyArgs *) args) ;
    else if (family == FAM_SAXPY_HEAVY)
        saxpy_heavy_kernel<S> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const SaxpyArgs *) args) ;
    else if (family == FAM_DOTG)
    {
        const int smem = 2 * DOTG_SLOTS * (int) sizeof (int32_t) ;
        static bool attr_set = false ;          // one flag per instantiation
        if (!attr_set)
        {
            cudaFuncSetAttribute (dotg_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) ;
            attr_set = true ;
        }
        dotg_kernel<S> <<<cfg.grid, cfg.block, smem, cfg.stream>>> (*(const DotGArgs *) args) ;
    }
    else if (family == FAM_DOTV)
        dotv_kernel<S> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const DotVArgs *) args) ;
    else if (family == FAM_DOTV_LONG)
        dotv_long_kernel<S> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const DotVArgs *) args) ;
    else if (family == FAM_SAXPYV)
        saxpyv_kernel<S> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const SaxpyVArgs *) args) ;
    else if (family == FAM_SAXPYV_LONG)
        saxpyv_long_kernel<S> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const SaxpyVArgs *) args) ;
    else if (family == FAM_SPMV)
        spmv_stream_kernel<S, false> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const SpmvArgs *) args) ;
    else if (family == FAM_SPMV_PRES)
        spmv_stream_kernel<S, true> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const SpmvArgs *) args) ;
    else
        dot_kernel<S> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const DotArgs *) args) ;
}

// T -> T semirings: 4 monoids (non-bool) or 4 boolean monoids (bool); T -> bool comparators:
// 4 boolean monoids.  The multiply operator is a uniform run-time switch (MULT = -1) except for
// the hot semirings listed in hot_semiring(), which get a compile-time operator.
template <class T>
inline bool launch_for_type (int family, int z_code, int add, int mult, const void *args,
    LaunchCfg cfg)
{
    constexpr bool Tbool = std::is_same<T, bool>::value ;
    if (z_code == GB200_BOOL)
    {
        switch (add)
        {
            case GB200_LOR  :
                if (Tbool && mult == GB200_LAND)
                    launch_family<Semiring<T, bool, GB200_LOR, GB200_LAND>> (family, args, cfg) ;
                else launch_family<Semiring<T, bool, GB200_LOR, -1>> (family, args, cfg) ;
                return true ;
            case GB200_LAND : launch_family<Semiring<T, bool, GB200_LAND, -1>> (family, args, cfg) ; return true ;
            case GB200_LXOR : launch_family<Semiring<T, bool, GB200_LXOR, -1>> (family, args, cfg) ; return true ;
            case GB200_EQ   : launch_family<Semiring<T, bool, GB200_EQ,   -1>> (family, args, cfg) ; return true ;
            default : return false ;
        }
    }
    if constexpr (!Tbool)
    {
        switch (add)
        {
            case GB200_MIN   :
                if (mult == GB200_PLUS)
                    launch_family<Semiring<T, T, GB200_MIN, GB200_PLUS>> (family, args, cfg) ;
                else launch_family<Semiring<T, T, GB200_MIN, -1>> (family, args, cfg) ;
                return true ;
            case GB200_MAX   : launch_family<Semiring<T, T, GB200_MAX,   -1>> (family, args, cfg) ; return true ;
            case GB200_PLUS  :
                if (mult == GB200_TIMES)
                    launch_family<Semiring<T, T, GB200_PLUS, GB200_TIMES>> (family, args, cfg) ;
                else launch_family<Semiring<T, T, GB200_PLUS, -1>> (family, args, cfg) ;
                return true ;
            case GB200_TIMES : launch_family<Semiring<T, T, GB200_TIMES, -1>> (family, args, cfg) ; return true ;
            default : return false ;
        }
    }
    return false ;
}

} // namespace gb200

#define GB200_INSTANTIATE_TYPE(NAME, T)                                                     \
    namespace gb200 {                                                                       \
    bool launch_##NAME (int family, int z_code, int add, int mult, const void *args,        \
        LaunchCfg cfg)                                                                      \
    { return launch_for_type<T> (family, z_code, add, mult, args, cfg) ; } }
