// kernels_vec.cuh -- the semiring-templated kernels of the vector multiplies (B is n-by-1), the
// shapes GrB_mxv / GrB_vxm hand to GB_AxB_parallel.
//
//   pull: w = A'*u  (do_adotb; reference Source/Template/GB_AxB_dot_nomask.c:20-83,
//         dot_compmask.c:20-126, dot_mask.c:33-159 with B->vdim == 1, inner loop dot_cij.c:102-148)
//   push: w = A*u   (saxpy with B->vdim == 1; reference Gustavson_nomask.c:66-159 / heap_mask.c)
//
// GPU design: a vector is cheap to hold densely, so u is expanded once into a value array plus a
// presence bitmap over [0,vlen) (or used in place when every entry is present), the mask becomes a
// bitmap, and both directions become one streaming pass over the traversed vectors of A:
//   pull = CSR SpMV over the semiring (one group of lanes per vector of A, gather from u),
//   push = SpMSpV (one warp per entry of u, monoid atomics into a dense accumulator + bitmap).
// Vectors of A longer than VEC_LONG entries are cut into segments handled by separate warps so
// that one hub vertex cannot serialise the step.
#pragma once
#include "common.cuh"
#include "semiring.cuh"

namespace gb200 {

constexpr int64_t VEC_LONG = 2048 ;         // longest run of entries handled by one lane group
constexpr int64_t VEC_HUGE = 32768 ;        // push: longer vectors of A are spread over the whole grid

struct VecItem { int64_t v ; int64_t p0, p1 ; } ;      // segment [p0,p1) of stored vector v of A

__device__ __forceinline__ bool bit_test (const uint32_t *__restrict__ bm, int64_t k)
{
    return (__ldg (bm + (k >> 5)) >> (k & 31)) & 1u ;
}

struct DotVArgs
{
    DMat A ;
    const void *bval ;              // u as a dense array of vlen values (garbage where absent)
    const uint32_t *bpres ;         // presence bitmap of u, or nullptr: every entry is present
    const uint32_t *mbits ;         // mask bitmap over the names of A's vectors, or nullptr
    int mask_comp ;
    void *vals ;                    // acc_t per stored vector of A
    uint8_t *flags ;                // 1 iff the vector produced an entry
    int G ;                         // lanes per vector
    const VecItem *items ;          // long vectors, cut into segments
    int64_t nitems ;
    unsigned long long *nmatch ;
    int mult_op ; int flip ;
} ;

// one group of G lanes per stored vector of A (vectors longer than VEC_LONG are left to
// dotv_long_kernel: here they only get their accumulator set to the identity)
template <class S>
__global__ void __launch_bounds__ (256)
dotv_kernel (DotVArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bv = (const T *) a.bval ;
    const int32_t *__restrict__ Ai = a.A.i ;
    acc_t *__restrict__ vals = (acc_t *) a.vals ;
    const int G = a.G ;
    const int gl = threadIdx.x & (G - 1) ;
    const int wl = threadIdx.x & 31 ;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(G - 1))) ;
    const int64_t gpb = blockDim.x / G ;
    const int64_t gstride = (int64_t) gridDim.x * gpb ;
    const int64_t nvec = a.A.nvec ;
    unsigned long long nm = 0 ;
    for (int64_t ia = (int64_t) blockIdx.x * gpb + threadIdx.x / G ; ia < nvec ; ia += gstride)
    {
        // every lane of a group sees the same ia, so group-wide votes below are safe
        const int64_t pa = __ldg (a.A.p + ia), pe = __ldg (a.A.p + ia + 1) ;
        bool live = (pe > pa) ;
        if (live && a.mbits)
        {
            const bool m = bit_test (a.mbits, dm_vecname (a.A, ia)) ;
            live = a.mask_comp ? !m : m ;
        }
        acc_t cij = Mon::identity () ;
        bool found = false ;
        const bool is_long = (pe - pa > VEC_LONG) ;
        if (live && !is_long)
        {
            for (int64_t base = pa ; base < pe ; base += G)
            {
                const int64_t p = base + gl ;
                if (p < pe)
                {
                    const int64_t k = __ldg (Ai + p) ;
                    if (a.bpres == nullptr || bit_test (a.bpres, k))
                    {
                        cij = Mon::combine (cij, sr.product (Ax [p], Bv [k])) ;
                        found = true ; nm++ ;
                    }
                }
                if (Mon::has_terminal ())
                {
                    // the terminal value is absorbing: the whole group can stop
                    if (__any_sync (gmask, found && Mon::is_terminal (cij))) break ;
                }
            }
        }
        const unsigned fm = __ballot_sync (gmask, found) ;
        for (int off = G >> 1 ; off > 0 ; off >>= 1)
        {
            const acc_t other = __shfl_down_sync (gmask, cij, off, G) ;
            if (gl + off < G) cij = Mon::combine (cij, other) ;
        }
        if (gl == 0)
        {
            if (live && is_long) { vals [ia] = Mon::identity () ; a.flags [ia] = 0 ; }
            else { a.flags [ia] = (fm != 0) ? 1 : 0 ; if (fm != 0) vals [ia] = cij ; }
        }
    }
    for (int off = 16 ; off > 0 ; off >>= 1) nm += __shfl_down_sync (0xffffffffu, nm, off) ;
    if (wl == 0 && nm) atomicAdd (a.nmatch, nm) ;
}

// one warp per segment of a long vector; partial results meet in the accumulator
template <class S>
__global__ void __launch_bounds__ (256)
dotv_long_kernel (DotVArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bv = (const T *) a.bval ;
    const int32_t *__restrict__ Ai = a.A.i ;
    acc_t *__restrict__ vals = (acc_t *) a.vals ;
    const int lane = threadIdx.x & 31 ;
    const int64_t wid = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
    const int64_t nw = ((int64_t) gridDim.x * blockDim.x) >> 5 ;
    unsigned long long nm = 0 ;
    for (int64_t it = wid ; it < a.nitems ; it += nw)
    {
        const VecItem item = a.items [it] ;
        if (a.mbits)
        {
            const bool m = bit_test (a.mbits, dm_vecname (a.A, item.v)) ;
            if (a.mask_comp ? m : !m) continue ;
        }
        acc_t cij = Mon::identity () ;
        bool found = false ;
        for (int64_t p = item.p0 + lane ; p < item.p1 ; p += 32)
        {
            const int64_t k = __ldg (Ai + p) ;
            if (a.bpres == nullptr || bit_test (a.bpres, k))
            {
                cij = Mon::combine (cij, sr.product (Ax [p], Bv [k])) ;
                found = true ; nm++ ;
                if (Mon::has_terminal () && Mon::is_terminal (cij)) break ;
            }
        }
        const unsigned fm = __ballot_sync (0xffffffffu, found) ;
        if (fm == 0) continue ;
        for (int off = 16 ; off > 0 ; off >>= 1)
            cij = Mon::combine (cij, __shfl_down_sync (0xffffffffu, cij, off)) ;
        if (lane == 0)
        {
            Mon::atomic_combine (vals + item.v, cij) ;
            a.flags [item.v] = 1 ;
        }
    }
    for (int off = 16 ; off > 0 ; off >>= 1) nm += __shfl_down_sync (0xffffffffu, nm, off) ;
    if (lane == 0 && nm) atomicAdd (a.nmatch, nm) ;
}

// ---------------------------------------------------------------------------------------------
// pull without a mask, streamed: w = A'*u as an entry-balanced SpMV over the semiring.
//
// dotv_kernel gives every vector of A one lane group, so on a power-law matrix most lanes of a
// group idle (short vectors) or one group drags on (hubs).  Here the ENTRIES of A are cut into tiles
// of SPMV_TILE consecutive entries, whatever vectors they belong to: a block streams its tile
// (indices and values fully coalesced, u gathered through L2), leaves the products in shared memory,
// and then reduces the vector segments that lie in the tile -- a thread per short segment, a warp per
// segment longer than 32.  A vector that lies inside one tile is stored directly; one that straddles
// tiles is combined with the monoid's atomic.  tile_row [t] (depends on A only, cached on the
// handle) is the stored vector that holds entry t * SPMV_TILE.
// ---------------------------------------------------------------------------------------------
// streaming load (evict-first): the entry stream of A must not push the gathered vector out of L2
template <class W> __device__ __forceinline__ W ld_stream (const W *ptr)
{
    if constexpr (sizeof (W) == 1)
    { const unsigned char r = __ldcs ((const unsigned char *) ptr) ; return *(const W *) &r ; }
    else if constexpr (sizeof (W) == 2)
    { const unsigned short r = __ldcs ((const unsigned short *) ptr) ; return *(const W *) &r ; }
    else if constexpr (sizeof (W) == 4)
    { const unsigned int r = __ldcs ((const unsigned int *) ptr) ; return *(const W *) &r ; }
    else
    { const unsigned long long r = __ldcs ((const unsigned long long *) ptr) ; return *(const W *) &r ; }
}

constexpr int SPMV_TILE = 2048 ;
constexpr int SPMV_THREADS = 256 ;
constexpr int SPMV_PER_THREAD = SPMV_TILE / SPMV_THREADS ;

struct SpmvArgs
{
    DMat A ;
    const void *bval ;              // u as a dense array of vlen values
    const uint32_t *bpres ;         // presence bitmap of u (HAS_PRES instantiation), else nullptr
    const int32_t *tile_row ;       // ntiles entries
    int64_t ntiles ;
    void *vals ;                    // acc_t per stored vector of A, pre-set to the identity
    uint8_t *flags ;                // pre-zeroed
    unsigned long long *nmatch ;
    int mult_op ; int flip ;
} ;

// OCC8: compiled for 8 resident blocks per SM (32 registers) instead of 6
template <class S, bool HAS_PRES, bool OCC8>
__global__ void __launch_bounds__ (SPMV_THREADS, OCC8 ? 8 : 6)
spmv_stream_kernel (SpmvArgs a)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    __shared__ acc_t sprod [SPMV_TILE] ;
    __shared__ uint32_t spres [SPMV_TILE / 32] ;
    __shared__ int32_t s_long [SPMV_TILE / 32] ;
    __shared__ int s_nlong ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Ax = (const T *) a.A.x ;
    const T *__restrict__ Bv = (const T *) a.bval ;
    const int32_t *__restrict__ Ai = a.A.i ;
    const int64_t *__restrict__ Ap = a.A.p ;
    acc_t *__restrict__ vals = (acc_t *) a.vals ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5 ;
    const int64_t nnz = a.A.nnz ;
    unsigned long long nm = 0 ;
    for (int64_t tile = blockIdx.x ; tile < a.ntiles ; tile += gridDim.x)
    {
        const int64_t e0 = tile * SPMV_TILE ;
        const int64_t e1 = (e0 + SPMV_TILE < nnz) ? (e0 + SPMV_TILE) : nnz ;
        // ---- phase 1: products of the tile into shared memory --------------------------------
        #pragma unroll
        for (int k = 0 ; k < SPMV_PER_THREAD ; k++)
        {
            const int idx = k * SPMV_THREADS + tid ;
            const int64_t p = e0 + idx ;
            bool f = false ;
            acc_t v = Mon::identity () ;
            if (p < e1)
            {
                const int64_t j = ld_stream (Ai + p) ;
                if (!HAS_PRES || bit_test (a.bpres, j))
                {
                    v = sr.product (ld_stream (Ax + p), Bv [j]) ;
                    f = true ;
                }
            }
            sprod [idx] = v ;
            if (HAS_PRES)
            {
                const unsigned m = __ballot_sync (0xffffffffu, f) ;
                if (lane == 0) spres [idx >> 5] = m ;
            }
            nm += f ? 1 : 0 ;
        }
        if (tid == 0) s_nlong = 0 ;
        __syncthreads () ;
        // ---- phase 2: the vector segments inside the tile ------------------------------------
        const int64_t r0 = a.tile_row [tile] ;
        const int64_t r_end = a.tile_row [tile + 1] ;       // ntiles + 1 entries
        for (int64_t r = r0 + tid ; r <= r_end ; r += SPMV_THREADS)
        {
            const int64_t ps = __ldg (Ap + r), pe = __ldg (Ap + r + 1) ;
            const int s = (int) (((ps > e0) ? ps : e0) - e0) ;
            const int e = (int) (((pe < e1) ? pe : e1) - e0) ;
            if (e <= s) continue ;
            if (e - s > 32) { s_long [atomicAdd (&s_nlong, 1)] = (int32_t) (r - r0) ; continue ; }
            acc_t acc = Mon::identity () ;
            bool found = false ;
            for (int q = s ; q < e ; q++)
            {
                if (!HAS_PRES || ((spres [q >> 5] >> (q & 31)) & 1u))
                {
                    const acc_t v = sprod [q] ;
                    acc = found ? Mon::combine (acc, v) : v ;
                    found = true ;
                }
            }
            if (found)
            {
                if (ps >= e0 && pe <= e1) vals [r] = acc ; else Mon::atomic_combine (vals + r, acc) ;
                a.flags [r] = 1 ;
            }
        }
        __syncthreads () ;
        const int nlong = s_nlong ;
        for (int k = warp ; k < nlong ; k += SPMV_THREADS / 32)
        {
            const int64_t r = r0 + s_long [k] ;
            const int64_t ps = __ldg (Ap + r), pe = __ldg (Ap + r + 1) ;
            const int s = (int) (((ps > e0) ? ps : e0) - e0) ;
            const int e = (int) (((pe < e1) ? pe : e1) - e0) ;
            acc_t acc = Mon::identity () ;
            bool found = false ;
            for (int q = s + lane ; q < e ; q += 32)
            {
                if (!HAS_PRES || ((spres [q >> 5] >> (q & 31)) & 1u))
                {
                    const acc_t v = sprod [q] ;
                    acc = found ? Mon::combine (acc, v) : v ;
                    found = true ;
                }
            }
            const unsigned fm = __ballot_sync (0xffffffffu, found) ;
            if (fm == 0) continue ;
            if (!found) acc = Mon::identity () ;
            for (int off = 16 ; off > 0 ; off >>= 1)
                acc = Mon::combine (acc, __shfl_down_sync (0xffffffffu, acc, off)) ;
            if (lane == 0)
            {
                if (ps >= e0 && pe <= e1) vals [r] = acc ; else Mon::atomic_combine (vals + r, acc) ;
                a.flags [r] = 1 ;
            }
        }
        __syncthreads () ;
    }
    for (int off = 16 ; off > 0 ; off >>= 1) nm += __shfl_down_sync (0xffffffffu, nm, off) ;
    if (lane == 0 && nm) atomicAdd (a.nmatch, nm) ;
}

// ---------------------------------------------------------------------------------------------
// push: w = A*u.  acc (vlen accumulators, pre-set to the identity) and pres (vlen bits, zeroed).
// ---------------------------------------------------------------------------------------------
struct SaxpyVArgs
{
    DMat A, B ;                     // B: the n-by-1 operand (its entries are B.i / B.x [0..B.nnz))
    const uint32_t *mbits ;         // mask bitmap over [0,vlen), or nullptr
    int mask_comp ;
    void *acc ;
    uint32_t *pres ;
    int32_t *longlist ;             // positions pb of entries of u whose vector of A is long
    int32_t *hugelist ;             // ... or huge (filled from the end of the same array)
    unsigned int *nlong ;           // nlong [0]: long, nlong [1]: huge
    unsigned long long *nflops ;
    int mult_op ; int flip ;
} ;

template <class S>
__device__ __forceinline__ void saxpyv_scatter (const SaxpyVArgs &a, const S &sr, int64_t p,
    typename S::T bk)
{
    using T = typename S::T ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    const int64_t i = __ldg (a.A.i + p) ;
    if (a.mbits)
    {
        const bool m = bit_test (a.mbits, i) ;
        if (a.mask_comp ? m : !m) return ;
    }
    acc_t *w = ((acc_t *) a.acc) + i ;
    const acc_t t = sr.product (((const T *) a.A.x) [p], bk) ;
    bool skip = false ;
    if constexpr (Mon::add == GB200_MIN || Mon::add == GB200_MAX || Mon::add == GB200_LOR
        || Mon::add == GB200_LAND)
    {
        // monotone, idempotent monoids: a product that cannot change the (possibly stale) value
        // read here cannot change the current one either; hub entries are hit by thousands of
        // products and this keeps them from serialising on one L2 atomic
        const acc_t old = *((volatile acc_t *) w) ;
        skip = Mon::memcmp_eq (Mon::combine (old, t), old) ;
    }
    if (!skip) Mon::atomic_combine (w, t) ;
    const uint32_t bit = 1u << (i & 31) ;
    uint32_t *pw = a.pres + (i >> 5) ;
    if (!(*((volatile uint32_t *) pw) & bit)) atomicOr (pw, bit) ;
}

// one warp per entry u(k)
template <class S>
__global__ void __launch_bounds__ (256)
saxpyv_kernel (SaxpyVArgs a)
{
    using T = typename S::T ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Bx = (const T *) a.B.x ;
    const int lane = threadIdx.x & 31 ;
    const int64_t wid = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
    const int64_t nw = ((int64_t) gridDim.x * blockDim.x) >> 5 ;
    unsigned long long nf = 0 ;
    for (int64_t pb = wid ; pb < a.B.nnz ; pb += nw)
    {
        const int64_t k = __ldg (a.B.i + pb) ;
        int64_t pa, pe ;
        if (!dm_lookup (a.A, k, pa, pe)) continue ;
        if (lane == 0) nf += (unsigned long long) (pe - pa) ;
        if (pe - pa > VEC_LONG)
        {
            if (lane == 0)
            {
                if (pe - pa > VEC_HUGE) a.hugelist [- (int64_t) atomicAdd (a.nlong + 1, 1u)] = (int32_t) pb ;
                else a.longlist [atomicAdd (a.nlong, 1u)] = (int32_t) pb ;
            }
            continue ;
        }
        const T bk = Bx [pb] ;
        for (int64_t p = pa + lane ; p < pe ; p += 32) saxpyv_scatter<S> (a, sr, p, bk) ;
    }
    if (lane == 0 && nf) atomicAdd (a.nflops, nf) ;
}

// the long vectors met by saxpyv_kernel: one block per long vector; every block takes a slice of
// each huge vector in turn
template <class S>
__global__ void __launch_bounds__ (256)
saxpyv_long_kernel (SaxpyVArgs a)
{
    using T = typename S::T ;
    const S sr (a.mult_op, a.flip != 0) ;
    const T *__restrict__ Bx = (const T *) a.B.x ;
    const unsigned int nl = a.nlong [0], nh = a.nlong [1] ;
    for (unsigned int q = blockIdx.x ; q < nl ; q += gridDim.x)
    {
        const int64_t pb = a.longlist [q] ;
        int64_t pa, pe ;
        dm_lookup (a.A, __ldg (a.B.i + pb), pa, pe) ;
        const T bk = Bx [pb] ;
        for (int64_t p = pa + threadIdx.x ; p < pe ; p += blockDim.x) saxpyv_scatter<S> (a, sr, p, bk) ;
    }
    for (unsigned int q = 0 ; q < nh ; q++)
    {
        const int64_t pb = a.hugelist [- (int64_t) q] ;
        int64_t pa, pe ;
        dm_lookup (a.A, __ldg (a.B.i + pb), pa, pe) ;
        const T bk = Bx [pb] ;
        for (int64_t p = pa + (int64_t) blockIdx.x * blockDim.x + threadIdx.x ; p < pe ;
            p += (int64_t) gridDim.x * blockDim.x) saxpyv_scatter<S> (a, sr, p, bk) ;
    }
}

// ---------------------------------------------------------------------------------------------
// reduce to a scalar: s = (+) of all entries of A over a built-in monoid -- the device half of
// GrB_reduce (reference Source/GB_reduce_to_scalar.c:107-270; SURVEY.md 8f row f3: the call that follows
// the triangle-counting multiply, Demo/Source/tricount.c:177).  Grid-stride partial results, a fixed
// shuffle / shared-memory tree per block, and the last block (ticket) combines the blocks' partials in
// block order: the result is deterministic for a given grid.
// ---------------------------------------------------------------------------------------------
struct ReduceArgs
{
    const void *x ;             // n values of the monoid's type
    int64_t n ;
    void *partial ;             // acc_t per block
    void *out ;                 // acc_t
    unsigned int *ticket ;      // zeroed before launch
} ;

template <class S>
__global__ void __launch_bounds__ (256)
reduce_kernel (ReduceArgs a)
{
    using Z = typename S::Z ; using acc_t = typename S::acc_t ; using Mon = typename S::Mon ;
    __shared__ acc_t s_part [8] ;
    __shared__ bool s_last ;
    const Z *__restrict__ x = (const Z *) a.x ;
    acc_t *__restrict__ partial = (acc_t *) a.partial ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5 ;
    acc_t acc = Mon::identity () ;
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < a.n ;
        t += (int64_t) gridDim.x * blockDim.x) acc = Mon::combine (acc, Mon::widen (x [t])) ;
    for (int off = 16 ; off > 0 ; off >>= 1) acc = Mon::combine (acc, __shfl_down_sync (0xffffffffu, acc, off)) ;
    if (lane == 0) s_part [warp] = acc ;
    __syncthreads () ;
    if (threadIdx.x == 0)
    {
        acc_t b = s_part [0] ;
        for (int w = 1 ; w < (int) (blockDim.x >> 5) ; w++) b = Mon::combine (b, s_part [w]) ;
        partial [blockIdx.x] = b ;
        __threadfence () ;
        s_last = (atomicAdd (a.ticket, 1u) == gridDim.x - 1) ;
    }
    __syncthreads () ;
    if (s_last && threadIdx.x == 0)
    {
        __threadfence () ;
        acc_t r = ((volatile acc_t *) partial) [0] ;
        for (unsigned int b = 1 ; b < gridDim.x ; b++) r = Mon::combine (r, ((volatile acc_t *) partial) [b]) ;
        *((acc_t *) a.out) = r ;
    }
}

} // namespace gb200
