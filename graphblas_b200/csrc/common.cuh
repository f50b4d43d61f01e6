// common.cuh -- device-side matrix view, vector lookup, small helpers shared by all kernels.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace gb200 {

// A device-resident sparse matrix as the kernels see it.  Vector pointers stay 64-bit (nnz can
// exceed 2^31); indices inside vectors are narrowed to 32 bits at upload (the library declines
// vlen or vdim >= 2^31), which halves the index traffic of every kernel.
struct DMat
{
    const int64_t *p ;      // nvec+1
    const int64_t *h ;      // nvec, or nullptr
    const int32_t *i ;      // nnz
    const void    *x ;      // nnz * type size
    int64_t vlen, vdim, nvec, nnz ;
    int hyper ;             // GB_IS_HYPER: is_hyper && nvec < vdim  (Source/GB.h:266-267)
    int type_code ;
    int iso ;               // 1: every stored value equals x[0] (a pattern-only matrix)
} ;

// Find vector k of A: returns [pa, pe).  Standard form: direct.  Hypersparse: binary search of
// the hyperlist (the role of GB_lookup, reference Source/GB.h:3396-3445).
__device__ __forceinline__ bool dm_lookup (const DMat &A, int64_t k, int64_t &pa, int64_t &pe)
{
    if (!A.hyper)
    {
        pa = __ldg (A.p + k) ; pe = __ldg (A.p + k + 1) ;
        return pe > pa ;
    }
    int64_t lo = 0, hi = A.nvec - 1 ;
    while (lo <= hi)
    {
        int64_t mid = (lo + hi) >> 1 ;
        int64_t hv = __ldg (A.h + mid) ;
        if (hv == k) { pa = __ldg (A.p + mid) ; pe = __ldg (A.p + mid + 1) ; return pe > pa ; }
        if (hv < k) lo = mid + 1 ; else hi = mid - 1 ;
    }
    pa = pe = 0 ;
    return false ;
}

// name of the kk-th stored vector
__device__ __forceinline__ int64_t dm_vecname (const DMat &A, int64_t kk)
{
    return A.hyper ? __ldg (A.h + kk) : kk ;
}

// position of `key` in the ascending list idx[lo..hi), or -1
__device__ __forceinline__ int64_t bsearch_i32 (const int32_t *__restrict__ idx, int64_t lo,
    int64_t hi, int32_t key)
{
    while (lo < hi)
    {
        int64_t mid = (lo + hi) >> 1 ;
        int32_t v = __ldg (idx + mid) ;
        if (v == key) return mid ;
        if (v < key) lo = mid + 1 ; else hi = mid ;
    }
    return -1 ;
}

// The products of one vector of B, A(:,k) (x) B(k,j) for every entry B(k,j) at positions [pb0, pb1), walked
// by the warps of the block as ONE flat index space: a warp takes a chunk of up to 32 entries of B, every
// lane looks up the vector of A of its entry (pointer loads of the whole chunk in flight together), a
// warp scan of the lengths lays the chunk's products end to end, and lane t of every round takes product
// t: f (p, pb) with p the position in A and pb the position in B.  An entry-after-entry walk costs a
// chain of dependent loads (B's index -> A's pointers -> A's entries) PER ENTRY and keeps as many lanes
// busy as A's vector is long (Erdos-Renyi, 8 per vector: 8 lanes, 8 chains one after another; measured
// 4.6 ms of a 8.0 ms multiply).  Must be called by all threads of the block; blockDim.x a multiple of 32.
// warp, nwarps: which of how many warps share the vector (a warp on its own: 0 of 1)
template <class F>
__device__ __forceinline__ void for_each_product_w (const DMat &A, const DMat &B, int64_t pb0, int64_t pb1,
    int warp, int nwarps, F &&f)
{
    constexpr unsigned FULL = 0xffffffffu ;
    constexpr int LONG = 64 ;       // a vector of A this long is walked by the whole warp on its own
    const int lane = threadIdx.x & 31 ;
    // entries of B per chunk: a lone warp takes up to 32 at a time; the warps of a larger block get about
    // four chunks each, dealt round-robin, so that one chunk of long vectors does not hold the block up
    int64_t cs = (nwarps == 1) ? 32 : (pb1 - pb0 + 4 * nwarps - 1) / (4 * nwarps) ;
    cs = (cs < 1) ? 1 : ((cs > 32) ? 32 : cs) ;
    for (int64_t c0 = pb0 + warp * cs ; c0 < pb1 ; c0 += nwarps * cs)
    {
        const int64_t pb = c0 + lane ;
        int64_t pa = 0, pe = 0 ;
        if (lane < cs && pb < pb1) { if (!dm_lookup (A, __ldg (B.i + pb), pa, pe)) { pa = 0 ; pe = 0 ; } }
        // long vectors of A: one after another, lanes striding (coalesced, no search)
        unsigned lm = __ballot_sync (FULL, pe - pa >= LONG) ;
        while (lm)
        {
            const int e = __ffs (lm) - 1 ;
            lm &= lm - 1 ;
            const int64_t pae = __shfl_sync (FULL, pa, e), pee = __shfl_sync (FULL, pe, e) ;
            for (int64_t p = pae + lane ; p < pee ; p += 32) f (p, c0 + e) ;
        }
        // the short ones end to end (fewer than 32 * LONG products: 32-bit offsets)
        const int len = (pe - pa >= LONG) ? 0 : (int) (pe - pa) ;
        int incl = len ;
        #pragma unroll
        for (int o = 1 ; o < 32 ; o <<= 1)
        {
            const int y = __shfl_up_sync (FULL, incl, o) ;
            if (lane >= o) incl += y ;
        }
        const int total = __shfl_sync (FULL, incl, 31) ;
        const int off = incl - len ;                        // non-decreasing over the lanes
        for (int t0 = 0 ; t0 < total ; t0 += 32)
        {
            const int t = t0 + lane ;
            int e = 0 ;                                     // the last entry whose products start at or before t
            #pragma unroll
            for (int sft = 16 ; sft > 0 ; sft >>= 1)
            {
                const int o = __shfl_sync (FULL, off, e + sft) ;
                if (o <= t) e += sft ;
            }
            const int oe = __shfl_sync (FULL, off, e) ;
            const int64_t pae = __shfl_sync (FULL, pa, e) ;
            if (t < total) f (pae + (t - oe), c0 + e) ;
        }
    }
}

template <class F>
__device__ __forceinline__ void for_each_product (const DMat &A, const DMat &B, int64_t pb0, int64_t pb1, F &&f)
{
    for_each_product_w (A, B, pb0, pb1, (int) (threadIdx.x >> 5), (int) (blockDim.x >> 5), f) ;
}

__device__ __forceinline__ uint32_t hash32 (uint32_t k) { return k * 0x9E3779B1u ; }

// one work item of a heavy column: B entries [pb0,pb1) of stored vector kk, workspace slot w
struct HeavyItem { int32_t kk ; int32_t w ; int64_t pb0 ; int64_t pb1 ; } ;

} // namespace gb200
