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
    // optional per-vector hash index (built lazily for the dot kernels): vectors longer than
    // VECHASH_MIN own an open-addressing table of 2^log slots (keys = indices, vals = offset of the
    // index inside the vector); hinfo[kk] = (table offset << 6) | log, or -1 for short vectors
    const int64_t *hinfo ;
    const int32_t *hkeys ;
    const int32_t *hofs ;
    // optional 64 KB Bloom filter per long vector (same vectors as the hash index), precomputed so
    // that a thread block can copy it into shared memory instead of rebuilding it: hbloom +
    // hbinfo[kk] * BLOOM_WORDS, hbinfo[kk] < 0 (or hbinfo == nullptr): none
    const int32_t *hbinfo ;
    const uint32_t *hbloom ;
    int iso ;               // 1: every stored value equals x[0] (a pattern-only matrix)
} ;

constexpr int BLOOM_WORDS = 16384 ;         // 2^19 bits
__device__ __forceinline__ uint32_t bloom_bit (uint32_t k) { return ((k ^ (k >> 15)) * 0x85EBCA6Bu) >> 13 ; }

constexpr int64_t VECHASH_MIN = 4096 ;     // == DOTG_CAP: shorter owners use shared memory

// position (in A.i / A.x) of index `key` in the vector [q0,q1) whose hash descriptor is `hi`, or -1
__device__ __forceinline__ int64_t vechash_probe (const DMat &A, int64_t hi, int64_t q0, int32_t key)
{
    const int lg = (int) (hi & 63) ;
    const int64_t off = hi >> 6 ;
    const uint32_t mask = (1u << lg) - 1u ;
    uint32_t h = ((uint32_t) key * 0x9E3779B1u) >> (32 - lg) ;
    while (true)
    {
        const int32_t k = __ldg (A.hkeys + off + h) ;
        if (k == key) return q0 + __ldg (A.hofs + off + h) ;
        if (k < 0) return -1 ;
        h = (h + 1) & mask ;
    }
}

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
__device__ __forceinline__ uint32_t hash32b (uint32_t k) { return (k ^ (k >> 15)) * 0x85EBCA6Bu ; }

// one work item of a heavy column: B entries [pb0,pb1) of stored vector kk, workspace slot w
struct HeavyItem { int32_t kk ; int32_t w ; int64_t pb0 ; int64_t pb1 ; } ;

} // namespace gb200
