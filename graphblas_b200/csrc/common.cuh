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

__device__ __forceinline__ uint32_t hash32 (uint32_t k) { return k * 0x9E3779B1u ; }

// one work item of a heavy column: B entries [pb0,pb1) of stored vector kk, workspace slot w
struct HeavyItem { int32_t kk ; int32_t w ; int64_t pb0 ; int64_t pb1 ; } ;

} // namespace gb200
