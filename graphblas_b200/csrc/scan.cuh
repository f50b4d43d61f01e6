// scan.cuh -- hand-written single-pass exclusive scan (decoupled look-back) and block helpers.
// Replaces GB_cumsum (reference Source/GB_cumsum.c:36-89) for Bflops, vector pointers of C and
// the stream compaction of masked results.
#pragma once
#include <type_traits>
#include <cstdint>
#include <cuda_runtime.h>

namespace gb200 {

__device__ __forceinline__ int64_t warp_sum_i64 (int64_t v)
{
    for (int off = 16 ; off > 0 ; off >>= 1) v += __shfl_down_sync (0xffffffffu, v, off) ;
    return v ;
}

// inclusive scan across a warp
__device__ __forceinline__ int64_t warp_incl_scan_i64 (int64_t v, int lane)
{
    for (int off = 1 ; off < 32 ; off <<= 1)
    {
        int64_t o = __shfl_up_sync (0xffffffffu, v, off) ;
        if (lane >= off) v += o ;
    }
    return v ;
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix and the block
// total.  `ws` must hold 33 int64 of shared memory.  blockDim.x <= 1024, a multiple of 32.
__device__ __forceinline__ int64_t block_excl_scan_i64 (int64_t v, int64_t *ws, int64_t &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5 ;
    int64_t incl = warp_incl_scan_i64 (v, lane) ;
    if (lane == 31) ws [warp] = incl ;
    __syncthreads () ;
    if (warp == 0)
    {
        int64_t w = (lane < nwarps) ? ws [lane] : 0 ;
        int64_t wi = warp_incl_scan_i64 (w, lane) ;
        if (lane < nwarps) ws [lane] = wi - w ;          // exclusive prefix of each warp
        if (lane == 31) ws [32] = wi ;                   // block total
    }
    __syncthreads () ;
    int64_t r = ws [warp] + incl - v ;
    total = ws [32] ;
    __syncthreads () ;
    return r ;
}

constexpr int SCAN_THREADS = 256 ;
constexpr int SCAN_ITEMS = 16 ;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS ;

struct ScanState
{
    unsigned int *ticket ;          // one counter, zeroed before launch
    int *flag ;                     // per tile: 0 = nothing, 1 = aggregate ready, 2 = inclusive ready
    int64_t *aggregate ;            // per tile
    int64_t *inclusive ;            // per tile
} ;

// out[t] = sum of in[0..t) for t = 0..n  (n+1 outputs; out[n] is the total).  Single pass with
// decoupled look-back.  A warp owns 32 * SCAN_ITEMS consecutive items, item q of lane l being
// q * 32 + l of them, so every load and store is coalesced; rows of 32 are scanned with shuffles.  The
// look-back over the predecessor tiles is done by a whole warp, 32 tiles per step.
template <class InT>
__global__ void __launch_bounds__ (SCAN_THREADS)
scan_kernel (const InT *__restrict__ in, int64_t *__restrict__ out, int64_t n, ScanState st)
{
    // the sum of one row of 32 small items fits 32 bits
    using row_t = typename std::conditional<(sizeof (InT) < 8), int32_t, int64_t>::type ;
    __shared__ int64_t ws [SCAN_THREADS / 32 + 1] ;
    __shared__ unsigned int s_tile ;
    __shared__ int64_t s_prefix ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5 ;
    if (threadIdx.x == 0) s_tile = atomicAdd (st.ticket, 1u) ;
    __syncthreads () ;
    const int64_t tile = s_tile ;
    const int64_t base = tile * SCAN_TILE + (int64_t) warp * (32 * SCAN_ITEMS) + lane ;
    int64_t ex [SCAN_ITEMS] ;       // exclusive prefix of every item inside the warp's chunk
    int64_t carry = 0 ;
    #pragma unroll
    for (int q = 0 ; q < SCAN_ITEMS ; q++)
    {
        const int64_t idx = base + q * 32 ;
        const row_t x = (idx < n) ? (row_t) in [idx] : (row_t) 0 ;
        row_t incl = x ;
        #pragma unroll
        for (int off = 1 ; off < 32 ; off <<= 1)
        {
            const row_t y = __shfl_up_sync (0xffffffffu, incl, off) ;
            if (lane >= off) incl += y ;
        }
        ex [q] = carry + (int64_t) (incl - x) ;
        carry += (int64_t) __shfl_sync (0xffffffffu, incl, 31) ;
    }
    if (lane == 0) ws [warp] = carry ;
    __syncthreads () ;
    if (warp == 0)
    {
        // exclusive prefix of the warps' totals, and the tile's total
        const int64_t w = (lane < SCAN_THREADS / 32) ? ws [lane] : 0 ;
        int64_t wi = w ;
        #pragma unroll
        for (int off = 1 ; off < 32 ; off <<= 1)
        {
            const int64_t y = __shfl_up_sync (0xffffffffu, wi, off) ;
            if (lane >= off) wi += y ;
        }
        const int64_t total = __shfl_sync (0xffffffffu, wi, 31) ;
        if (lane < SCAN_THREADS / 32) ws [lane] = wi - w ;
        int64_t prefix = 0 ;
        if (tile == 0)
        {
            if (lane == 0)
            {
                ((volatile int64_t *) st.inclusive) [0] = total ;
                __threadfence () ;
                ((volatile int *) st.flag) [0] = 2 ;
            }
        }
        else
        {
            if (lane == 0)
            {
                ((volatile int64_t *) st.aggregate) [tile] = total ;
                __threadfence () ;
                ((volatile int *) st.flag) [tile] = 1 ;
            }
            for (int64_t t = tile - 1 ; ; t -= 32)
            {
                const int64_t idx = t - lane ;
                int f = 2 ;                         // before tile 0: an inclusive prefix of zero
                int64_t val = 0 ;
                if (idx >= 0)
                {
                    while ((f = ((volatile int *) st.flag) [idx]) == 0) { }
                    __threadfence () ;
                    val = (f == 2) ? ((volatile int64_t *) st.inclusive) [idx]
                                   : ((volatile int64_t *) st.aggregate) [idx] ;
                }
                // the nearest predecessor with an inclusive prefix ends the walk
                const unsigned done = __ballot_sync (0xffffffffu, f == 2) ;
                const int last = done ? (__ffs (done) - 1) : 31 ;
                int64_t part = (lane <= last) ? val : 0 ;
                #pragma unroll
                for (int off = 16 ; off > 0 ; off >>= 1) part += __shfl_down_sync (0xffffffffu, part, off) ;
                prefix += __shfl_sync (0xffffffffu, part, 0) ;
                if (done) break ;
            }
            if (lane == 0)
            {
                ((volatile int64_t *) st.inclusive) [tile] = prefix + total ;
                __threadfence () ;
                ((volatile int *) st.flag) [tile] = 2 ;
            }
        }
        if (lane == 0) s_prefix = prefix ;
    }
    __syncthreads () ;
    const int64_t add = s_prefix + ws [warp] ;
    #pragma unroll
    for (int q = 0 ; q < SCAN_ITEMS ; q++)
    {
        const int64_t idx = base + q * 32 ;
        if (idx <= n) out [idx] = add + ex [q] ;
    }
}

} // namespace gb200
