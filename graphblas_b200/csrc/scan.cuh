// scan.cuh -- hand-written single-pass exclusive scan (decoupled look-back) and block helpers.
// Replaces GB_cumsum (reference Source/GB_cumsum.c:36-89) for Bflops, vector pointers of C and
// the stream compaction of masked results.
#pragma once
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
constexpr int SCAN_ITEMS = 8 ;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS ;

struct ScanState
{
    unsigned int *ticket ;          // one counter, zeroed before launch
    int *flag ;                     // per tile: 0 = nothing, 1 = aggregate ready, 2 = inclusive ready
    int64_t *aggregate ;            // per tile
    int64_t *inclusive ;            // per tile
} ;

// out[t] = sum of in[0..t) for t = 0..n  (n+1 outputs; out[n] is the total)
template <class InT>
__global__ void __launch_bounds__ (SCAN_THREADS)
scan_kernel (const InT *__restrict__ in, int64_t *__restrict__ out, int64_t n, ScanState st)
{
    __shared__ int64_t ws [33] ;
    __shared__ unsigned int s_tile ;
    __shared__ int64_t s_prefix ;
    if (threadIdx.x == 0) s_tile = atomicAdd (st.ticket, 1u) ;
    __syncthreads () ;
    const int64_t tile = s_tile ;
    const int64_t base = tile * SCAN_TILE + (int64_t) threadIdx.x * SCAN_ITEMS ;
    int64_t v [SCAN_ITEMS] ;
    int64_t tsum = 0 ;
    #pragma unroll
    for (int q = 0 ; q < SCAN_ITEMS ; q++)
    {
        const int64_t idx = base + q ;
        v [q] = (idx < n) ? (int64_t) in [idx] : 0 ;
        tsum += v [q] ;
    }
    int64_t total ;
    int64_t texcl = block_excl_scan_i64 (tsum, ws, total) ;
    if (threadIdx.x == 0)
    {
        int64_t prefix = 0 ;
        if (tile == 0)
        {
            ((volatile int64_t *) st.inclusive) [0] = total ;
            __threadfence () ;
            ((volatile int *) st.flag) [0] = 2 ;
        }
        else
        {
            ((volatile int64_t *) st.aggregate) [tile] = total ;
            __threadfence () ;
            ((volatile int *) st.flag) [tile] = 1 ;
            for (int64_t t = tile - 1 ; t >= 0 ; t--)
            {
                int f ;
                while ((f = ((volatile int *) st.flag) [t]) == 0) { }
                __threadfence () ;
                if (f == 2) { prefix += ((volatile int64_t *) st.inclusive) [t] ; break ; }
                prefix += ((volatile int64_t *) st.aggregate) [t] ;
            }
            ((volatile int64_t *) st.inclusive) [tile] = prefix + total ;
            __threadfence () ;
            ((volatile int *) st.flag) [tile] = 2 ;
        }
        s_prefix = prefix ;
    }
    __syncthreads () ;
    int64_t run = s_prefix + texcl ;
    #pragma unroll
    for (int q = 0 ; q < SCAN_ITEMS ; q++)
    {
        const int64_t idx = base + q ;
        if (idx <= n) out [idx] = run ;
        run += v [q] ;
    }
}

} // namespace gb200
