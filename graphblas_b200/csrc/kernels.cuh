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
#include "scan.cuh"
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
    // fused symbolic-fill + numeric (saxpy_hash_kernel)
    int hash_log ;              // the shared-memory table has 2^hash_log slots
    int32_t *Ci_out ;           // pattern of C, written by the kernel (ascending in every vector)
} ;

// ---------------------------------------------------------------------------------------------
// saxpy, fused symbolic fill + numeric phase for vectors of C whose pattern fits a shared-memory hash
// table (the north star's "shared-memory hash accumulator"; reference behaviour:
// Source/Template/GB_AxB_Gustavson_symbolic.c:187-233 + GB_AxB_Gustavson_nomask.c:91-158 -- there a
// dense Work/Mark pair of O(vlen) plus GB_qsort_1, here a table of (row, accumulator) slots).
// One thread block per vector of B.  Every product A(i,k) (x) B(k,j) claims the slot of row i with one
// shared atomicCAS and is combined into the slot's accumulator with the monoid's shared atomic; the
// occupied slots are then compacted, sorted by row with a bitonic network in shared memory and
// written out coalesced: C's pattern and values leave the SM exactly once, A's entries are read once.
// When the index range is small (nwords > 0: a vlen-bit bitmap fits shared memory next to the table)
// the rows are also marked in the bitmap and come out of it ascending by prefix popcount: no sort.
// The exact number of entries of every vector (lp) comes from the count pass.
// ---------------------------------------------------------------------------------------------
template <class S>
__global__ void saxpy_hash_kernel (SaxpyArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    extern __shared__ __align__ (16) unsigned char hash_raw [] ;
    const int LOG = a.hash_log ;
    const int size = 1 << LOG ;
    acc_t *vals = (acc_t *) hash_raw ;                          // size accumulators
    uint64_t *comp = (uint64_t *) (vals + size) ;               // size / 2 (row << 32 | slot) pairs
    int32_t *keys = (int32_t *) (comp + size / 2) ;             // size rows, -1 = free
    uint32_t *bm = (uint32_t *) (keys + size) ;                 // nwords words (small index ranges only)
    const int nwords = (int) a.nwords ;
    __shared__ int s_n ;
    __shared__ int64_t s_ws [33] ;
    const uint32_t mask = (uint32_t) size - 1u ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bx = (const T *) a.B.x ;
    acc_t *__restrict__ acc = (acc_t *) a.acc ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5 ;
    for (int64_t c = blockIdx.x ; c < a.ncols ; c += gridDim.x)
    {
        const int64_t kk = a.cols [c] ;
        const int64_t base = a.lp [kk] ;
        if (a.lp [kk+1] <= base) continue ;                     // block-uniform
        for (int t = threadIdx.x ; t < size ; t += blockDim.x) { keys [t] = -1 ; vals [t] = Mon::identity () ; }
        for (int t = threadIdx.x ; t < nwords ; t += blockDim.x) bm [t] = 0u ;
        if (threadIdx.x == 0) s_n = 0 ;
        __syncthreads () ;
        const int64_t pb0 = a.B.p [kk], pb1 = a.B.p [kk+1] ;
        for_each_product (a.A, a.B, pb0, pb1, [&] (int64_t p, int64_t pb)
        {
            const int32_t i = __ldg (a.A.i + p) ;
            const acc_t prod = sr.product (Ax [p], Bx [pb]) ;
            uint32_t h = (hash32 ((uint32_t) i) >> (32 - LOG)) & mask ;
            while (true)
            {
                const int32_t old = atomicCAS (keys + h, -1, i) ;
                if (old == -1 || old == i)
                {
                    Mon::atomic_combine (vals + h, prod) ;
                    if (old == -1 && nwords > 0) atomicOr (bm + (((uint32_t) i) >> 5), 1u << (i & 31)) ;
                    break ;
                }
                h = (h + 1) & mask ;
            }
        }) ;
        __syncthreads () ;
        if (nwords > 0)
        {
            // small index range: the rows were also marked in a bitmap, which hands them out ascending
            // (rank = prefix popcount) -- no sort; the value of a row is looked up in the table
            int64_t run = base ;
            for (int t0 = 0 ; t0 < nwords ; t0 += blockDim.x)
            {
                const int t = t0 + threadIdx.x ;
                uint32_t word = (t < nwords) ? bm [t] : 0u ;
                int64_t total ;
                int64_t q = run + block_excl_scan_i64 (__popc (word), s_ws, total) ;
                while (word)
                {
                    const int32_t key = t * 32 + (__ffs (word) - 1) ;
                    word &= word - 1 ;
                    uint32_t h = (hash32 ((uint32_t) key) >> (32 - LOG)) & mask ;
                    while (keys [h] != key) h = (h + 1) & mask ;
                    a.Ci_out [q] = key ;
                    acc [q] = vals [h] ;
                    q++ ;
                }
                run += total ;
            }
            __syncthreads () ;
            continue ;
        }
        if (blockDim.x == 32)
        {
            // one warp per vector: the occupied slots are counted with votes, and up to 64 of them are
            // sorted in registers (two per lane, a bitonic network of shuffles: no shared-memory round
            // trips, no barriers)
            constexpr unsigned FULL = 0xffffffffu ;
            int n = 0 ;
            for (int t0 = 0 ; t0 < size ; t0 += 32)
            {
                const int32_t key = keys [t0 + lane] ;
                const unsigned occ = __ballot_sync (FULL, key >= 0) ;
                if (key >= 0) comp [n + __popc (occ & ((1u << lane) - 1u))] = ((uint64_t) (uint32_t) key << 32) | (uint32_t) (t0 + lane) ;
                n += __popc (occ) ;
            }
            __syncwarp () ;
            if (n <= 64)
            {
                uint64_t x0 = (lane < n) ? comp [lane] : ~0ULL ;
                uint64_t x1 = (lane + 32 < n) ? comp [lane + 32] : ~0ULL ;
                #pragma unroll
                for (int k = 2 ; k <= 64 ; k <<= 1)
                {
                    #pragma unroll
                    for (int j = k >> 1 ; j > 0 ; j >>= 1)
                    {
                        if (j == 32)
                        {
                            if (x0 > x1) { const uint64_t y = x0 ; x0 = x1 ; x1 = y ; }
                        }
                        else
                        {
                            const uint64_t y0 = __shfl_xor_sync (FULL, x0, j), y1 = __shfl_xor_sync (FULL, x1, j) ;
                            const bool low = ((lane & j) == 0) ;            // this lane holds the pair's lower index
                            const bool asc0 = ((lane & k) == 0), asc1 = (((lane + 32) & k) == 0) ;
                            // the lower index keeps the minimum of an ascending pair
                            x0 = ((x0 < y0) == (low == asc0)) ? x0 : y0 ;
                            x1 = ((x1 < y1) == (low == asc1)) ? x1 : y1 ;
                        }
                    }
                }
                if (lane < n) { a.Ci_out [base + lane] = (int32_t) (x0 >> 32) ; acc [base + lane] = vals [(uint32_t) x0] ; }
                if (lane + 32 < n) { a.Ci_out [base + lane + 32] = (int32_t) (x1 >> 32) ; acc [base + lane + 32] = vals [(uint32_t) x1] ; }
                __syncwarp () ;
                continue ;
            }
            if (lane == 0) s_n = n ;
        }
        else
        {
            for (int t = threadIdx.x ; t < size ; t += blockDim.x)
            {
                const int32_t key = keys [t] ;
                if (key >= 0) comp [atomicAdd (&s_n, 1)] = ((uint64_t) (uint32_t) key << 32) | (uint32_t) t ;
            }
        }
        __syncthreads () ;
        const int n = s_n ;
        int n2 = 1 ; while (n2 < n) n2 <<= 1 ;
        for (int t = n + threadIdx.x ; t < n2 ; t += blockDim.x) comp [t] = ~0ULL ;
        __syncthreads () ;
        for (int k = 2 ; k <= n2 ; k <<= 1)
        {
            for (int j = k >> 1 ; j > 0 ; j >>= 1)
            {
                for (int t = threadIdx.x ; t < n2 ; t += blockDim.x)
                {
                    const int ixj = t ^ j ;
                    if (ixj > t)
                    {
                        const uint64_t x = comp [t], y = comp [ixj] ;
                        const bool asc = ((t & k) == 0) ;
                        if ((x > y) == asc) { comp [t] = y ; comp [ixj] = x ; }
                    }
                }
                __syncthreads () ;
            }
        }
        for (int t = threadIdx.x ; t < n ; t += blockDim.x)
        {
            const uint64_t v = comp [t] ;
            a.Ci_out [base + t] = (int32_t) (v >> 32) ;
            acc [base + t] = vals [(uint32_t) v] ;
        }
        __syncthreads () ;
    }
}

// ---------------------------------------------------------------------------------------------
// The same for vectors of C with at most 128 entries when no bitmap is used: a WARP per vector, several
// independent warps per block.  With one 32-thread block per vector an SM holds 32 warps (the limit on
// resident blocks), each running the chain B's pointers -> B's entries -> A's pointers -> A's entries ->
// table -> sort -> write for one vector at a time; three warps per block and 16 blocks per SM give it
// 48.  Everything is warp-synchronous (no block barrier); a warp's table is its own 4 KB of the block's
// shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int HASHW_WARPS = 3 ;
constexpr int HASHW_LOG = 8 ;

template <class S> __host__ __device__ constexpr int hashw_bytes ()
{
    return (int) ((sizeof (typename S::acc_t) + 8) << HASHW_LOG) ;      // accumulators + keys + sort pairs
}

template <class S>
__global__ void __launch_bounds__ (32 * HASHW_WARPS, 16)
saxpy_hash_warp_kernel (SaxpyArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    constexpr unsigned FULL = 0xffffffffu ;
    constexpr int LOG = HASHW_LOG, size = 1 << LOG ;
    constexpr uint32_t mask = (uint32_t) size - 1u ;
    extern __shared__ __align__ (16) unsigned char hashw_raw [] ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5 ;
    unsigned char *mine = hashw_raw + warp * hashw_bytes<S> () ;
    acc_t *vals = (acc_t *) mine ;                              // size accumulators
    uint64_t *comp = (uint64_t *) (vals + size) ;               // size / 2 (row << 32 | slot) pairs
    int32_t *keys = (int32_t *) (comp + size / 2) ;             // size rows, -1 = free
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bx = (const T *) a.B.x ;
    acc_t *__restrict__ acc = (acc_t *) a.acc ;
    for (int64_t c = (int64_t) blockIdx.x * HASHW_WARPS + warp ; c < a.ncols ; c += (int64_t) gridDim.x * HASHW_WARPS)
    {
        const int64_t kk = a.cols [c] ;
        const int64_t base = a.lp [kk] ;
        if (a.lp [kk+1] <= base) continue ;                     // warp-uniform
        for (int t = lane ; t < size ; t += 32) { keys [t] = -1 ; vals [t] = Mon::identity () ; }
        __syncwarp () ;
        for_each_product_w (a.A, a.B, a.B.p [kk], a.B.p [kk+1], 0, 1, [&] (int64_t p, int64_t pb)
        {
            const int32_t i = __ldg (a.A.i + p) ;
            const acc_t prod = sr.product (Ax [p], Bx [pb]) ;
            uint32_t h = (hash32 ((uint32_t) i) >> (32 - LOG)) & mask ;
            while (true)
            {
                const int32_t old = atomicCAS (keys + h, -1, i) ;
                if (old == -1 || old == i) { Mon::atomic_combine (vals + h, prod) ; break ; }
                h = (h + 1) & mask ;
            }
        }) ;
        __syncwarp () ;
        // the occupied slots, counted with votes
        int n = 0 ;
        for (int t0 = 0 ; t0 < size ; t0 += 32)
        {
            const int32_t key = keys [t0 + lane] ;
            const unsigned occ = __ballot_sync (FULL, key >= 0) ;
            if (key >= 0) comp [n + __popc (occ & ((1u << lane) - 1u))] = ((uint64_t) (uint32_t) key << 32) | (uint32_t) (t0 + lane) ;
            n += __popc (occ) ;
        }
        __syncwarp () ;
        if (n <= 64)
        {
            // two per lane, sorted by a bitonic network of shuffles
            uint64_t x0 = (lane < n) ? comp [lane] : ~0ULL ;
            uint64_t x1 = (lane + 32 < n) ? comp [lane + 32] : ~0ULL ;
            #pragma unroll
            for (int k = 2 ; k <= 64 ; k <<= 1)
            {
                #pragma unroll
                for (int j = k >> 1 ; j > 0 ; j >>= 1)
                {
                    if (j == 32)
                    {
                        if (x0 > x1) { const uint64_t y = x0 ; x0 = x1 ; x1 = y ; }
                    }
                    else
                    {
                        const uint64_t y0 = __shfl_xor_sync (FULL, x0, j), y1 = __shfl_xor_sync (FULL, x1, j) ;
                        const bool low = ((lane & j) == 0) ;
                        const bool asc0 = ((lane & k) == 0), asc1 = (((lane + 32) & k) == 0) ;
                        x0 = ((x0 < y0) == (low == asc0)) ? x0 : y0 ;
                        x1 = ((x1 < y1) == (low == asc1)) ? x1 : y1 ;
                    }
                }
            }
            if (lane < n) { a.Ci_out [base + lane] = (int32_t) (x0 >> 32) ; acc [base + lane] = vals [(uint32_t) x0] ; }
            if (lane + 32 < n) { a.Ci_out [base + lane + 32] = (int32_t) (x1 >> 32) ; acc [base + lane + 32] = vals [(uint32_t) x1] ; }
        }
        else
        {
            // 65 .. 128 entries: the bitonic network over the warp's shared memory
            for (int t = n + lane ; t < 128 ; t += 32) comp [t] = ~0ULL ;
            __syncwarp () ;
            for (int k = 2 ; k <= 128 ; k <<= 1)
            {
                for (int j = k >> 1 ; j > 0 ; j >>= 1)
                {
                    for (int t = lane ; t < 128 ; t += 32)
                    {
                        const int ixj = t ^ j ;
                        if (ixj > t)
                        {
                            const uint64_t x = comp [t], y = comp [ixj] ;
                            const bool asc = ((t & k) == 0) ;
                            if ((x > y) == asc) { comp [t] = y ; comp [ixj] = x ; }
                        }
                    }
                    __syncwarp () ;
                }
            }
            for (int t = lane ; t < n ; t += 32)
            {
                const uint64_t v = comp [t] ;
                a.Ci_out [base + t] = (int32_t) (v >> 32) ;
                acc [base + t] = vals [(uint32_t) v] ;
            }
        }
        __syncwarp () ;
    }
}

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
        for_each_product (a.A, a.B, pb0, pb1, [&] (int64_t p, int64_t pb)
        {
            const int32_t i = __ldg (a.A.i + p) ;
            const int64_t slot = bsearch_i32 (a.li, l0, l1, i) ;
            if (slot < 0) return ;                  // only possible when masked
            Mon::atomic_combine (acc + slot, sr.product (Ax [p], Bx [pb])) ;
            if (a.flags) a.flags [slot] = 1 ;
        }) ;
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
    const int32_t *plist ;      // DOT_MASK: the entries of M to compute (nullptr: all npairs of them)
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
        const int64_t t = gid0 + itn * gstride ;
        bool live = (t < a.npairs) ;
        const int64_t e = (live && a.plist) ? (int64_t) a.plist [t] : t ;
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
                        // sparse case: the first product is copied, later ones combined (dot_cij.c:29-45)
                        const acc_t prod = sr.product (Ax [p], Bx [l]) ;
                        cij = found ? Mon::combine (cij, prod) : prod ;
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
                        const acc_t prod = sr.product (Ax [l], Bx [p]) ;
                        cij = found ? Mon::combine (cij, prod) : prod ;
                        found = true ; nm++ ;
                        if (Mon::has_terminal () && Mon::is_terminal (cij)) break ;
                    }
                }
            }
        }
        // combine the G partial results (a fixed tree: deterministic for floating point)
        const unsigned fm0 = __ballot_sync (0xffffffffu, found) ;
        {
            // only lanes that hold a product take part: the reference copies the first product and
            // combines the later ones (GB_AxB_dot_cij.c:29-45), so a pair whose only products are NaN is
            // NaN under MIN / MAX, where combining with the identity would give +-Inf
            const int wl = threadIdx.x & 31 ;
            unsigned fm = fm0 ;
            for (int off = G >> 1 ; off > 0 ; off >>= 1)
            {
                acc_t other = __shfl_down_sync (0xffffffffu, cij, off, G) ;
                const bool of = (gl + off < G) && ((fm >> (wl + off)) & 1u) ;
                if (of) cij = ((fm >> wl) & 1u) ? Mon::combine (cij, other) : other ;
                // lane l now also holds what lane l + off held (inside its own group only)
                fm |= (gl + off < G) ? (((fm >> (wl + off)) & 1u) << wl) : 0u ;
                fm = __reduce_or_sync (0xffffffffu, fm & (1u << wl)) ;
            }
        }
        const unsigned fm = fm0 ;
        if (t < a.npairs && gl == 0)
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
// power-law graph the walked lengths sum to ~8.6x the number of matches (~5x once the walks are trimmed
// to the owner's index range, engine_dot.cu), so a probe must cost a
// handful of instructions and never leave the SM: pairs are grouped by their owner vector, a thread
// block loads the owner into a shared-memory CUCKOO table (two tables, two hash functions, total
// load <= 3/8) once and then serves every task of its work item from it.  A cuckoo lookup is
// exactly two shared-memory loads and two compares -- no probe loop, so no divergence -- which lets
// the walk be unrolled DOTG_U-fold with all its loads in flight.  Owners longer than the table's
// capacity ("hubs") are loaded segment by segment (both lists are sorted, so every task keeps a
// cursor and each walked index is probed against exactly one segment) by the HUB instantiation, in
// which a LANE, not a warp, walks a task (dotg_lanes).  orient 0: owner = B(:,j); orient 1:
// owner = A(:,i) (the mask entries regrouped by i).  A TASK is one pair, or one DOTG_SEG-long
// segment of a pair whose walked list is longer than that (pieces are combined with the monoid's
// atomic); the warps of a block pull tasks from a shared counter, so one long pair cannot stall a
// block.  ISO = both operands hold one repeated value (a pattern-only matrix, e.g. an adjacency
// matrix of ones): the product is a constant, no value is loaded, and a task reduces to counting
// its matches.  Pairs whose owner is shorter than DOTG_SMALL never get here (dot_kernel does them).
// ---------------------------------------------------------------------------------------------
constexpr int DOTG_SMEM = 64 * 1024 ;       // table bytes per block
constexpr int DOTG_SEG = 1024 ;             // longest walk of one task
constexpr int DOTG_THREADS = 512 ;
constexpr int DOTG_U = 4 ;                  // walk unrolling: 32 * DOTG_U indices per warp iteration
constexpr int DOTG_HUB_TASKS = 16384 ;      // most tasks of one HUB work item (one 16-bit cursor each)
constexpr int DOTG_SMALL = 32 ;             // owners shorter than this: dot_kernel, one lane group per pair
constexpr int DOTG_MAXIT = 48 ;             // longest eviction chain before a table is rebuilt

// owner entries per table load: the two tables together are filled to 3/8 at most
__host__ __device__ constexpr int dotg_cap (bool iso) { return (DOTG_SMEM / (iso ? 4 : 8)) * 3 / 8 ; }

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
    unsigned int *failed ;              // set to 1 if an owner's cuckoo tables could not be built: the
                                        // host then recomputes the pairs with the table-free dot_kernel
    int64_t bm_bits ;                   // dotr_kernel<BITMAP>: indices per bitmap part (a multiple of 32)
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

// c (+) c (+) ... (+) c, n >= 1 times, by doubling (any associative monoid; exact for the integer,
// boolean and MIN/MAX monoids)
template <class Mon> __device__ __forceinline__ typename Mon::acc_t iso_fold (typename Mon::acc_t c, uint32_t n)
{
    using acc_t = typename Mon::acc_t ;
    if constexpr (Mon::add == GB200_MIN || Mon::add == GB200_MAX || Mon::add == GB200_LOR
        || Mon::add == GB200_LAND) return c ;                          // idempotent
    else if constexpr (Mon::add == GB200_PLUS && !std::is_floating_point<acc_t>::value)
        return wrap_mul<acc_t> (c, (acc_t) n) ;
    else
    {
        acc_t r = c, base = c ;
        bool have = false ;
        while (n)
        {
            if (n & 1u) { r = have ? Mon::combine (r, base) : base ; have = true ; }
            n >>= 1 ;
            if (n) base = Mon::combine (base, base) ;
        }
        return r ;
    }
}

// what the task loop of one (item, segment) needs
template <class S> struct DotGSeg
{
    const DotTask *tasks ;      // of this item
    int ntask ;
    const int32_t *Wi ;         // walked matrix: indices, values
    const typename S::T *Wx ;
    const typename S::T *Ox ;   // owner values of this segment
    typename S::acc_t *vals ;
    uint8_t *flags ;
    int32_t vhi ;               // indices above vhi belong to a later segment
    int NS, sh ;                // table geometry and the two multipliers
    uint32_t c1, c2 ;
    bool orient ;
    typename S::acc_t ciso ;
} ;

// one cuckoo lookup: exactly two shared-memory loads, no loop (tab2 = the second table)
template <bool ISO, bool DENSE, class slot_t>
__device__ __forceinline__ bool dotg_probe (const slot_t *__restrict__ tab, const slot_t *__restrict__ tab2,
    uint32_t kq, int sh, uint32_t c1, uint32_t c2, uint32_t &pos)
{
    constexpr uint32_t NOKEY = 0xFFFFFFFEu ;
    if constexpr (DENSE) { pos = kq ; return (kq != NOKEY) ; }
    else if constexpr (ISO)
    {
        const uint32_t e1 = tab [(kq * c1) >> sh] ;
        const uint32_t e2 = tab2 [(kq * c2) >> sh] ;
        return (e1 == kq) | (e2 == kq) ;
    }
    else
    {
        const uint64_t e1 = tab [(kq * c1) >> sh] ;
        const uint64_t e2 = tab2 [(kq * c2) >> sh] ;
        const bool h1 = ((uint32_t) e1 == kq), h2 = ((uint32_t) e2 == kq) ;
        pos = (uint32_t) ((h1 ? e1 : e2) >> 32) ;
        return h1 | h2 ;
    }
}

// eight consecutive 32-bit indices with ONE 32-byte load (LDG.E.ENL2.256 on sm_100); p is 32-byte aligned
__device__ __forceinline__ void ldg256 (const int32_t *p, int32_t (&k) [8])
{
#ifdef GB200_HOST_EMULATION             // tools/emu_kernels.py runs this source on the host
    for (int c = 0 ; c < 8 ; c++) k [c] = p [c] ;
#else
    asm volatile ("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r" (k [0]), "=r" (k [1]), "=r" (k [2]), "=r" (k [3]), "=r" (k [4]),
          "=r" (k [5]), "=r" (k [6]), "=r" (k [7]) : "l" (p)) ;
#endif
}

// HUB items (owner longer than one table load): every LANE pulls tasks from the shared counter and
// walks, by itself, the part of its task that falls into the current segment -- eight indices (two
// 16-byte loads) per step, the cursor of every task kept in shared memory between segments.  There
// is no cross-lane reduction and the per-visit overhead is paid by one lane, not by a warp; hub items
// hold thousands of tasks, so the lanes stay busy.  The loop is a state machine so that the 32 lanes
// of a warp reconverge every step whatever the lengths of their tasks.
template <class S, bool ISO, class slot_t, bool LD256>
__device__ __forceinline__ void dotg_lanes (const S &sr, const DotGSeg<S> &g, const slot_t *tab,
    int *s_next, uint16_t *s_cur, unsigned long long &nm)
{
    // LD256: a chunk is one 32-byte load (LDG.E.ENL2.256, sm_100) instead of two 16-byte ones -- every
    // lane reads its own list, so a load costs one L1 wavefront per lane whatever its width (measured
    // on the pattern-only kernel, tri scale 22: 41.7 -> 40.5 ms; the valued kernel keeps 16-byte loads)
    constexpr int AL = LD256 ? 7 : 3 ;
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    constexpr uint32_t NOKEY = 0xFFFFFFFEu ;
    const int sh = g.sh ;
    const uint32_t c1 = g.c1, c2 = g.c2 ;
    const slot_t *__restrict__ tab2 = tab + g.NS ;
    int t = -1 ;                        // the task this lane is walking
    bool more = true ;                  // the counter may still hold tasks
    int32_t e = 0 ;
    const int32_t *wp = nullptr ;       // the walked list
    const T *wx = nullptr ;
    int pbeg = 0, len = 0, q = 0 ;      // first index to look at, end, next aligned chunk
    uint32_t cnt = 0 ;
    acc_t cij = Mon::identity () ;
    bool found = false ;
    while (true)
    {
        if (t < 0 && more)
        {
            const int c = atomicAdd (s_next, 1) ;
            if (c >= g.ntask) more = false ;
            else
            {
                const int cur = (int) s_cur [c] ;
                const DotTask d = g.tasks [c] ;
                const int ln = (d.len < 0) ? -d.len : d.len ;
                if (cur < ln)
                {
                    t = c ; e = d.e ; len = ln ; pbeg = cur ;
                    wp = g.Wi + d.w0 ; wx = g.Wx + d.w0 ;
                    // chunks are 16-byte (32-byte) aligned in the index array (its base is 256-byte
                    // aligned and its allocation a multiple of 512 bytes)
                    q = cur - (int) ((d.w0 + cur) & AL) ;
                    cnt = 0 ; found = false ; cij = Mon::identity () ;
                }
            }
        }
        if (!__any_sync (0xffffffffu, t >= 0 || more)) break ;
        if (t >= 0)
        {
            int32_t kk [8] ;
            if constexpr (LD256) ldg256 (wp + q, kk) ;
            else
            {
                const int4 ka = __ldg ((const int4 *) (wp + q)) ;
                int4 kb = make_int4 (0, 0, 0, 0) ;
                if (q + 4 < len) kb = __ldg ((const int4 *) (wp + q + 4)) ;
                kk [0] = ka.x ; kk [1] = ka.y ; kk [2] = ka.z ; kk [3] = ka.w ;
                kk [4] = kb.x ; kk [5] = kb.y ; kk [6] = kb.z ; kk [7] = kb.w ;
            }
            // Indices of the chunk outside [pbeg,len) belong to other lists and are not probed.  An
            // index above vhi, or one left behind by an earlier segment, needs no test: the table
            // holds the owner's indices of this segment only, so it simply misses.
            const unsigned span = (unsigned) (len - pbeg) ;
            #pragma unroll
            for (int c = 0 ; c < 8 ; c++)
            {
                const int p = q + c ;
                const uint32_t kq = ((unsigned) (p - pbeg) < span) ? (uint32_t) kk [c] : NOKEY ;
                uint32_t pos = 0 ;
                const bool hit = dotg_probe<ISO, false, slot_t> (tab, tab2, kq, sh, c1, c2, pos) ;
                if constexpr (ISO) cnt += hit ? 1u : 0u ;
                else if (hit)
                {
                    const T ov = g.Ox [pos], wv = wx [p] ;
                    const acc_t prod = g.orient ? sr.product (ov, wv) : sr.product (wv, ov) ;
                    cij = found ? Mon::combine (cij, prod) : prod ;
                    found = true ; cnt++ ;
                }
            }
            // the lists are sorted: the segment ends inside this chunk iff its last index is above vhi
            const int jl = (len - 1 - q < 7) ? (len - 1 - q) : 7 ;
            int32_t klast = kk [0] ;
            #pragma unroll
            for (int c = 1 ; c < 8 ; c++) klast = (jl >= c) ? kk [c] : klast ;
            bool stop = false ;
            int newcur = len ;
            if (klast > g.vhi)
            {
                stop = true ;
                newcur = (q > pbeg) ? q : pbeg ;
                #pragma unroll
                for (int c = 0 ; c < 8 ; c++)
                    if ((unsigned) (q + c - pbeg) < span && kk [c] <= g.vhi) newcur = q + c + 1 ;
            }
            q += 8 ;
            bool done = stop || (q >= len) ;
            if (Mon::has_terminal () && !ISO)
                if (found && Mon::is_terminal (cij)) { done = true ; newcur = len ; }
            if (done)
            {
                s_cur [t] = (uint16_t) newcur ;
                if (cnt)
                {
                    if constexpr (ISO) cij = iso_fold<Mon> (g.ciso, cnt) ;
                    Mon::atomic_combine (g.vals + e, cij) ;     // one piece per segment
                    g.flags [e] = 1 ;
                    nm += cnt ;
                }
                t = -1 ;
            }
        }
    }
}

// Regular items (the owner fits one table load): the warps of the block pull the tasks of the item
// and walk them against the table in `tab`, DOTG_U x 32 indices per iteration with all loads in
// flight.  DENSE: the owner holds every index, so there is no table.
template <class S, bool ISO, bool DENSE, class slot_t>
__device__ __forceinline__ void dotg_walk (const S &sr, const DotGSeg<S> &g, const slot_t *tab,
    int *s_next, unsigned long long &nm)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    constexpr uint32_t NOKEY = 0xFFFFFFFEu ;            // a walked index that must not be probed
    const int lane = threadIdx.x & 31 ;
    const int sh = g.sh ;
    const uint32_t c1 = g.c1, c2 = g.c2 ;
    const slot_t *__restrict__ tab2 = tab + g.NS ;
    int tk = 0 ;
    if (lane == 0) tk = atomicAdd (s_next, 1) ;
    tk = __shfl_sync (0xffffffffu, tk, 0) ;
    DotTask task ;
    if (tk < g.ntask) task = g.tasks [tk] ;
    while (tk < g.ntask)
    {
        // claim the next task and fetch its descriptor before working on this one
        int tkn = 0 ;
        if (lane == 0) tkn = atomicAdd (s_next, 1) ;
        tkn = __shfl_sync (0xffffffffu, tkn, 0) ;
        DotTask tnext ;
        if (tkn < g.ntask) tnext = g.tasks [tkn] ;

        const bool split = (task.len < 0) ;
        const int len = split ? -task.len : task.len ;
        const int32_t *__restrict__ wp = g.Wi + task.w0 ;
        acc_t cij = Mon::identity () ;
        bool found = false ;
        uint32_t cnt = 0 ;
        for (int p0 = 0 ; p0 < len ; p0 += 32 * DOTG_U)
        {
            uint32_t k [DOTG_U] ;
            #pragma unroll
            for (int u = 0 ; u < DOTG_U ; u++)
            {
                const int p = p0 + 32 * u + lane ;
                k [u] = (p < len) ? (uint32_t) __ldg (wp + p) : NOKEY ;
            }
            #pragma unroll
            for (int u = 0 ; u < DOTG_U ; u++)
            {
                if (u > 0 && p0 + 32 * u >= len) continue ;     // warp-uniform
                const int p = p0 + 32 * u + lane ;
                uint32_t pos = 0 ;
                const bool hit = dotg_probe<ISO, DENSE, slot_t> (tab, tab2, k [u], sh, c1, c2, pos) ;
                if constexpr (ISO) { if (hit) cnt++ ; }
                else if (hit)
                {
                    const T ov = g.Ox [pos], wv = g.Wx [task.w0 + p] ;
                    const acc_t prod = g.orient ? sr.product (ov, wv) : sr.product (wv, ov) ;
                    cij = found ? Mon::combine (cij, prod) : prod ;
                    found = true ; cnt++ ;
                }
            }
            if (Mon::has_terminal () && !ISO)
            {
                // the terminal value is absorbing: once any lane holds it the pair is decided
                if (__any_sync (0xffffffffu, found && Mon::is_terminal (cij))) break ;
            }
        }
        // ---- combine the lanes' partial results -------------------------------------------------
        bool any ;
        if constexpr (ISO)
        {
            cnt = __reduce_add_sync (0xffffffffu, cnt) ;
            any = (cnt != 0) ;
            if (any) cij = iso_fold<Mon> (g.ciso, cnt) ;
            if (lane == 0) nm += cnt ;
        }
        else
        {
            any = (__ballot_sync (0xffffffffu, found) != 0) ;
            nm += cnt ;
            if (any)
            {
                // only lanes that hold a product take part (see dot_kernel)
                unsigned fm = __ballot_sync (0xffffffffu, found) ;
                for (int off = 16 ; off > 0 ; off >>= 1)
                {
                    const acc_t other = __shfl_down_sync (0xffffffffu, cij, off) ;
                    const bool of = (lane + off < 32) && ((fm >> (lane + off)) & 1u) ;
                    if (of) cij = ((fm >> lane) & 1u) ? Mon::combine (cij, other) : other ;
                    fm |= (fm >> off) ;
                }
                // a dense owner: the reference starts from the identity there (dot_cij.c:110,125,141)
                if constexpr (DENSE) cij = Mon::combine (Mon::identity (), cij) ;
            }
        }
        if (any && lane == 0)
        {
            if (split) Mon::atomic_combine (g.vals + task.e, cij) ;
            else g.vals [task.e] = cij ;
            g.flags [task.e] = 1 ;
        }
        tk = tkn ; task = tnext ;
    }
}

// HUB = false: items whose owner fits one table load (or is dense); HUB = true: items of longer owners
template <class S, bool ISO, bool HUB>
__global__ void __launch_bounds__ (DOTG_THREADS, (ISO && !HUB) ? 3 : 2)
dotg_kernel (DotGArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    using slot_t = typename std::conditional<ISO, uint32_t, uint64_t>::type ;
    extern __shared__ __align__ (16) unsigned char dotg_raw [] ;
    slot_t *tab = (slot_t *) dotg_raw ;
    __shared__ int s_next, s_fail, s_skip ;
    __shared__ unsigned long long s_item ;
    __shared__ uint16_t s_cur [HUB ? DOTG_HUB_TASKS : 1] ;
    constexpr int CAP = dotg_cap (ISO) ;
    constexpr slot_t EMPTY = (slot_t) ~(slot_t) 0 ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bx = (const T *) a.B.x ;
    const int lane = threadIdx.x & 31 ;
    const bool orient = (a.orient != 0) ;
    const DMat &O = orient ? a.A : a.B ;        // owner matrix (probed)
    const DMat &W = orient ? a.B : a.A ;        // walked matrix
    const T *__restrict__ Oxb = orient ? Ax : Bx ;
    const int64_t vlen = a.A.vlen ;
    DotGSeg<S> g ;
    g.Wi = W.i ; g.Wx = orient ? Bx : Ax ;
    g.vals = (acc_t *) a.vals ; g.flags = a.flags ; g.orient = orient ;
    g.ciso = Mon::identity () ;
    if (ISO) g.ciso = sr.product (Ax [0], Bx [0]) ;
    unsigned long long nm = 0 ;
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
        if (!orient) ko = dm_vecpos (a.B, dm_vecname (a.M, item.owner)) ;
        const int64_t o0 = __ldg (O.p + ko), o1 = __ldg (O.p + ko + 1) ;
        const int64_t olen = o1 - o0 ;
        const bool dense = (olen == vlen) ;
        const int nseg = (!HUB || dense) ? 1 : (int) ((olen + CAP - 1) / CAP) ;
        g.tasks = a.tasks + item.e0 ;
        g.ntask = (int) (item.e1 - item.e0) ;
        if (HUB) for (int t = threadIdx.x ; t < g.ntask ; t += blockDim.x) s_cur [t] = 0 ;
        if (threadIdx.x == 0) s_skip = 0 ;
        for (int seg = 0 ; seg < nseg ; seg++)
        {
            const int64_t s0 = o0 + (int64_t) seg * CAP ;
            const int64_t s1 = (s0 + CAP < o1) ? (s0 + CAP) : o1 ;
            const int slen = (int) (s1 - s0) ;
            g.vhi = (seg == nseg - 1) ? INT32_MAX : __ldg (O.i + s1 - 1) ;
            g.Ox = Oxb + s0 ;
            int lg = 5 ;                        // 2^lg slots per table: total load between 3/16 and 3/8
            while (3 * (1 << lg) < 4 * slen) lg++ ;
            const int NS = 1 << lg, sh = 32 - lg ;
            uint32_t c1 = 0x9E3779B1u, c2 = 0x85EBCA6Bu ;
            if (!dense)
            {
                for (int attempt = 0 ; ; attempt++)
                {
                    __syncthreads () ;                  // the table's previous users are done
                    for (int t = threadIdx.x ; t < 2 * NS ; t += blockDim.x) tab [t] = EMPTY ;
                    if (threadIdx.x == 0) s_fail = 0 ;
                    __syncthreads () ;
                    for (int q = threadIdx.x ; q < slen ; q += blockDim.x)
                    {
                        slot_t cur ;
                        if constexpr (ISO) cur = (uint32_t) __ldg (O.i + s0 + q) ;
                        else cur = ((uint64_t) (uint32_t) q << 32) | (uint32_t) __ldg (O.i + s0 + q) ;
                        int which = 0, n = 0 ;
                        #pragma unroll 1
                        for ( ; n < DOTG_MAXIT ; n++)
                        {
                            const uint32_t k = (uint32_t) cur ;
                            const uint32_t loc = which ? (NS + ((k * c2) >> sh)) : ((k * c1) >> sh) ;
                            if constexpr (ISO) cur = atomicExch (tab + loc, cur) ;
                            else cur = atomicExch ((unsigned long long *) tab + loc, (unsigned long long) cur) ;
                            if (cur == EMPTY) break ;
                            which ^= 1 ;                // the evicted entry moves to its other table
                        }
                        if (n == DOTG_MAXIT) s_fail = 1 ;
                    }
                    __syncthreads () ;
                    if (!s_fail) break ;
                    if (attempt >= 30)
                    {
                        // never seen.  No trap (a trap poisons the context of the host process): the
                        // item is abandoned, the flag tells the host to recompute with dot_kernel
                        if (threadIdx.x == 0) { *a.failed = 1u ; s_skip = 1 ; }
                        break ;
                    }
                    c1 = (c1 * 0x01000193u + 0xFE94F82Au) | 1u ;
                    c2 = (c2 * 0x01000193u + 0x4A8BE922u) | 1u ;
                }
            }
            else __syncthreads () ;
            if (threadIdx.x == 0) s_next = 0 ;
            __syncthreads () ;
            if (s_skip) break ;
            g.NS = NS ; g.sh = sh ; g.c1 = c1 ; g.c2 = c2 ;
            if constexpr (HUB)
            {
                if (dense) dotg_walk<S, ISO, true, slot_t> (sr, g, tab, &s_next, nm) ;
                else dotg_lanes<S, ISO, slot_t, ISO> (sr, g, tab, &s_next, s_cur, nm) ;
            }
            else
            {
                if (dense) dotg_walk<S, ISO, true, slot_t> (sr, g, tab, &s_next, nm) ;
                else dotg_walk<S, ISO, false, slot_t> (sr, g, tab, &s_next, nm) ;
            }
        }
    }
    for (int off = 16 ; off > 0 ; off >>= 1) nm += __shfl_down_sync (0xffffffffu, nm, off) ;
    if (lane == 0 && nm) atomicAdd (a.nmatch, nm) ;
}

} // namespace gb200

#include "kernels_dotr.cuh"

namespace gb200 {

// ---------------------------------------------------------------------------------------------
// launchers, one set per (xy type); defined in inst_*.cu through GB200_INSTANTIATE_TYPE
// ---------------------------------------------------------------------------------------------
enum { FAM_SAXPY_LIGHT = 0, FAM_SAXPY_HEAVY = 1, FAM_DOT = 2, FAM_DOTG = 3, FAM_DOTV = 4,
    FAM_DOTV_LONG = 5, FAM_SAXPYV = 6, FAM_SAXPYV_LONG = 7, FAM_SPMV = 8, FAM_SPMV_PRES = 9,
    FAM_DOTG_ISO = 10, FAM_SPMV_OCC8 = 11, FAM_DOTG_HUB = 12, FAM_DOTG_HUB_ISO = 13,
    FAM_DOTR = 14, FAM_DOTR_ISO = 15, FAM_DOTR_BM = 16, FAM_DOTR_BM_ISO = 17, FAM_DOTR_WARP = 18,
    FAM_DOTR_WARP_ISO = 19, FAM_SAXPY_HASH = 20, FAM_REDUCE = 21, FAM_SAXPY_HASH_WARP = 22 } ;

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
    else if (family == FAM_REDUCE)
        reduce_kernel<S> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const ReduceArgs *) args) ;
    else if (family == FAM_SAXPY_HASH)
    {
        // per slot: accumulator + row + half a (row, slot) pair of the sort
        const SaxpyArgs &sa = *(const SaxpyArgs *) args ;
        const size_t smem = ((sizeof (typename S::acc_t) + 8) << sa.hash_log) + (size_t) sa.nwords * 4 ;
        static size_t attr_smem = 48 * 1024 ;       // one per instantiation
        if (smem > attr_smem)
        {
            cudaFuncSetAttribute (saxpy_hash_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) ;
            attr_smem = smem ;
        }
        saxpy_hash_kernel<S> <<<cfg.grid, cfg.block, smem, cfg.stream>>> (sa) ;
    }
    else if (family == FAM_SAXPY_HASH_WARP)
    {
        static bool attr_set = false ;          // one flag per instantiation
        if (!attr_set)
        {
            cudaFuncSetAttribute (saxpy_hash_warp_kernel<S>, cudaFuncAttributePreferredSharedMemoryCarveout, 100) ;
            attr_set = true ;
        }
        saxpy_hash_warp_kernel<S> <<<cfg.grid, cfg.block, HASHW_WARPS * hashw_bytes<S> (), cfg.stream>>> (*(const SaxpyArgs *) args) ;
    }
    else if (family == FAM_DOTG || family == FAM_DOTG_ISO || family == FAM_DOTG_HUB
        || family == FAM_DOTG_HUB_ISO)
    {
        static bool attr_set = false ;          // one flag per instantiation
        if (!attr_set)
        {
            cudaFuncSetAttribute (dotg_kernel<S, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DOTG_SMEM) ;
            cudaFuncSetAttribute (dotg_kernel<S, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DOTG_SMEM) ;
            cudaFuncSetAttribute (dotg_kernel<S, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DOTG_SMEM) ;
            cudaFuncSetAttribute (dotg_kernel<S, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DOTG_SMEM) ;
            attr_set = true ;
        }
        const DotGArgs &ga = *(const DotGArgs *) args ;
        if (family == FAM_DOTG_ISO)
            dotg_kernel<S, true, false> <<<cfg.grid, cfg.block, DOTG_SMEM, cfg.stream>>> (ga) ;
        else if (family == FAM_DOTG)
            dotg_kernel<S, false, false> <<<cfg.grid, cfg.block, DOTG_SMEM, cfg.stream>>> (ga) ;
        else if (family == FAM_DOTG_HUB_ISO)
            dotg_kernel<S, true, true> <<<cfg.grid, cfg.block, DOTG_SMEM, cfg.stream>>> (ga) ;
        else
            dotg_kernel<S, false, true> <<<cfg.grid, cfg.block, DOTG_SMEM, cfg.stream>>> (ga) ;
    }
    else if (family == FAM_DOTR || family == FAM_DOTR_ISO || family == FAM_DOTR_BM
        || family == FAM_DOTR_BM_ISO)
    {
        static bool attr_set = false ;          // one flag per instantiation
        if (!attr_set)
        {
            cudaFuncSetAttribute (dotr_kernel<S, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DOTG_SMEM + (DOTR_THREADS / 32) * DOTR_HIT_BYTES) ;
            cudaFuncSetAttribute (dotr_kernel<S, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DOTG_SMEM) ;
            cudaFuncSetAttribute (dotr_kernel<S, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DOTR_BM_SMEM + (DOTR_BM_THREADS / 32) * DOTR_HIT_BYTES) ;
            cudaFuncSetAttribute (dotr_kernel<S, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DOTR_BM_SMEM) ;
            attr_set = true ;
        }
        const DotGArgs &ga = *(const DotGArgs *) args ;
        if (family == FAM_DOTR_ISO)
            dotr_kernel<S, true, false> <<<cfg.grid, cfg.block, DOTG_SMEM, cfg.stream>>> (ga) ;
        else if (family == FAM_DOTR)
            dotr_kernel<S, false, false> <<<cfg.grid, cfg.block, DOTG_SMEM + (DOTR_THREADS / 32) * DOTR_HIT_BYTES, cfg.stream>>> (ga) ;
        else if (family == FAM_DOTR_BM_ISO)
            dotr_kernel<S, true, true> <<<cfg.grid, cfg.block, DOTR_BM_SMEM, cfg.stream>>> (ga) ;
        else
            dotr_kernel<S, false, true> <<<cfg.grid, cfg.block, DOTR_BM_SMEM + (DOTR_BM_THREADS / 32) * DOTR_HIT_BYTES, cfg.stream>>> (ga) ;
    }
    else if (family == FAM_DOTR_WARP || family == FAM_DOTR_WARP_ISO)
    {
        static bool attr_set = false ;          // one flag per instantiation
        if (!attr_set)
        {
            cudaFuncSetAttribute (dotr_warp_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DOTG_SMEM + (DOTR_THREADS / 32) * DOTR_HIT_BYTES) ;
            cudaFuncSetAttribute (dotr_warp_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DOTG_SMEM) ;
            attr_set = true ;
        }
        const DotGArgs &ga = *(const DotGArgs *) args ;
        if (family == FAM_DOTR_WARP_ISO)
            dotr_warp_kernel<S, true> <<<cfg.grid, cfg.block, DOTG_SMEM, cfg.stream>>> (ga) ;
        else
            dotr_warp_kernel<S, false> <<<cfg.grid, cfg.block, DOTG_SMEM + (DOTR_THREADS / 32) * DOTR_HIT_BYTES, cfg.stream>>> (ga) ;
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
        spmv_stream_kernel<S, false, false> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const SpmvArgs *) args) ;
    else if (family == FAM_SPMV_PRES)
        spmv_stream_kernel<S, true, false> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const SpmvArgs *) args) ;
    else if (family == FAM_SPMV_OCC8)
        spmv_stream_kernel<S, false, true> <<<cfg.grid, cfg.block, 0, cfg.stream>>> (*(const SpmvArgs *) args) ;
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
