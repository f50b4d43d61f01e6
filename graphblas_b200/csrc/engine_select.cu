// engine_select.cu -- GxB_select with the built-in operators on the device (the
// step that builds L = tril (A,-1) and U = triu (A,1) in front of the triangle-counting multiply,
// reference Demo/Program/tri_demo.c:80,95 -> Source/GB_select.c:219-323).
//
// Reference behaviour restated, in the CSC-agnostic terms of GB_select.c (j = name of the vector,
// i = index inside it, kk already negated and TRIL/TRIU already swapped by the caller when the matrix
// is held by row, GB_select.c:150-170):
//     TRIL     keep (j - i) <= kk          GB_select.c:226-244
//     TRIU     keep (j - i) >= kk          :249-262
//     DIAG     keep (j - i) == kk          :267-292
//     OFFDIAG  keep (j - i) != kk          :297-310
//     NONZERO  keep x != 0 (any byte set)  :315-328
// T has A's type and dimensions and A's hypersparsity (a hypersparse T lists only its non-empty
// vectors, GB_jappend); entries keep their order, so every vector stays ascending.
// GPU: one flag per entry, the library's single-pass scan, one gather.
#include "engine.cuh"
#include "scan.cuh"
#include "semiring.cuh"

namespace gb200 {

static inline int sel_grid (int64_t n, int per_sm)
{
    int64_t g = (n + 255) / 256, cap = (int64_t) ctx ().sm_count * per_sm ;
    if (g > cap) g = cap ;
    if (g < 1) g = 1 ;
    return (int) g ;
}

// a warp takes 32 consecutive vectors: short ones are flagged by their lane, long ones by the warp
__global__ void select_flag_kernel (DMat A, int op, int64_t kk, int tsz, uint8_t *__restrict__ keep)
{
    const int lane = threadIdx.x & 31 ;
    const int64_t wid = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
    const int64_t nw = ((int64_t) gridDim.x * blockDim.x) >> 5 ;
    const unsigned char *__restrict__ Ax = (const unsigned char *) A.x ;
    auto test = [&] (int64_t j, int64_t e) -> uint8_t
    {
        const int64_t d = j - (int64_t) __ldg (A.i + e) ;
        if (op == GB200_SELECT_TRIL) return d <= kk ;
        if (op == GB200_SELECT_TRIU) return d >= kk ;
        if (op == GB200_SELECT_DIAG) return d == kk ;
        if (op == GB200_SELECT_OFFDIAG) return d != kk ;
        bool nz = false ;
        for (int b = 0 ; b < tsz ; b++) nz = nz || (Ax [e * tsz + b] != 0) ;
        return nz ;
    } ;
    for (int64_t v0 = wid * 32 ; v0 < A.nvec ; v0 += nw * 32)
    {
        const int64_t v = v0 + lane ;
        int64_t e0 = 0, e1 = 0, j = 0 ;
        if (v < A.nvec) { e0 = __ldg (A.p + v) ; e1 = __ldg (A.p + v + 1) ; j = dm_vecname (A, v) ; }
        const bool is_long = (e1 - e0 > 64) ;
        if (!is_long) for (int64_t e = e0 ; e < e1 ; e++) keep [e] = test (j, e) ;
        unsigned todo = __ballot_sync (0xffffffffu, is_long) ;
        while (todo)
        {
            const int src = __ffs (todo) - 1 ;
            todo &= todo - 1 ;
            const int64_t s0 = __shfl_sync (0xffffffffu, e0, src), s1 = __shfl_sync (0xffffffffu, e1, src) ;
            const int64_t sj = __shfl_sync (0xffffffffu, j, src) ;
            for (int64_t e = s0 + lane ; e < s1 ; e += 32) keep [e] = test (sj, e) ;
        }
    }
}

__global__ void select_gather_kernel (const uint8_t *__restrict__ keep, const int64_t *__restrict__ pos,
    int64_t nnz, const int32_t *__restrict__ Ai, const unsigned char *__restrict__ Ax, int tsz,
    int32_t *__restrict__ Ti, unsigned char *__restrict__ Tx)
{
    for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; e < nnz ;
        e += (int64_t) gridDim.x * blockDim.x)
    {
        if (!keep [e]) continue ;
        const int64_t q = pos [e] ;
        Ti [q] = Ai [e] ;
        if (tsz == 8) ((uint64_t *) Tx) [q] = ((const uint64_t *) Ax) [e] ;
        else if (tsz == 4) ((uint32_t *) Tx) [q] = ((const uint32_t *) Ax) [e] ;
        else if (tsz == 2) ((uint16_t *) Tx) [q] = ((const uint16_t *) Ax) [e] ;
        else Tx [q] = Ax [e] ;
    }
}

// cum [v] = entries kept in the vectors before v
__global__ void select_cum_kernel (const int64_t *__restrict__ p, const int64_t *__restrict__ pos,
    int64_t nvec, int64_t *__restrict__ cum)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t <= nvec ;
        t += (int64_t) gridDim.x * blockDim.x) cum [t] = pos [p [t]] ;
}

} // namespace gb200

using namespace gb200 ;

extern "C" {
#pragma GCC visibility push(default)

gb200_status gb200_select_device (gb200_result *out, gb200_dmatrix Ad, int select_op, int64_t k)
{
    if (out == NULL || Ad == NULL || select_op < GB200_SELECT_TRIL || select_op > GB200_SELECT_NONZERO)
        return GB200_INVALID ;
    *out = NULL ;
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    const DMat &A = Ad->v ;
    const int64_t nnz = A.nnz ;
    const int tsz = type_size (A.type_code) ;
    gb200_result_s *R = new (std::nothrow) gb200_result_s () ;
    if (R == NULL) return GB200_OUT_OF_MEMORY ;
    memset (&R->info, 0, sizeof (R->info)) ;
    auto body = [&] () -> gb200_status
    {
        cudaEventRecord (c.ev0, c.stream) ;
        DevBuf keep, pos, cum, Ti, Tx ;
        GB200_TRY (keep.alloc (nnz > 0 ? nnz : 1)) ;
        GB200_TRY (pos.alloc ((nnz + 1) * sizeof (int64_t))) ;
        GB200_TRY (cum.alloc ((A.nvec + 1) * sizeof (int64_t))) ;
        if (nnz > 0)
        {
            select_flag_kernel <<<sel_grid ((A.nvec + 31) / 32 * 32, 16), 256, 0, c.stream>>> (A, select_op, k,
                tsz, keep.as<uint8_t> ()) ;
            count_launch () ;
        }
        GB200_TRY (scan_u8 (keep.as<uint8_t> (), pos.as<int64_t> (), nnz)) ;
        int64_t tnz = 0 ;
        GB200_TRY (read_i64 (pos.as<int64_t> () + nnz, &tnz)) ;
        select_cum_kernel <<<sel_grid (A.nvec + 1, 8), 256, 0, c.stream>>> (A.p, pos.as<int64_t> (), A.nvec,
            cum.as<int64_t> ()) ;
        count_launch () ;
        GB200_TRY (Ti.alloc ((tnz > 0 ? tnz : 1) * sizeof (int32_t))) ;
        GB200_TRY (Tx.alloc ((size_t) (tnz > 0 ? tnz : 1) * tsz)) ;
        if (tnz > 0)
        {
            select_gather_kernel <<<sel_grid (nnz, 16), 256, 0, c.stream>>> (keep.as<uint8_t> (),
                pos.as<int64_t> (), nnz, A.i, (const unsigned char *) A.x, tsz, Ti.as<int32_t> (),
                (unsigned char *) Tx.ptr) ;
            count_launch () ;
        }
        GB200_CUDA (cudaGetLastError ()) ;
        R->info.type_code = A.type_code ;
        R->info.method_used = 0 ; R->info.mask_applied = 0 ; R->info.flops = nnz ;
        GB200_TRY (assemble (R, A.nvec, A.hyper ? A.h : nullptr, A.hyper != 0, cum, Ti, Tx, tnz,
            Ad->is_hyper_flag != 0, A.vlen, A.vdim)) ;
        cudaEventRecord (c.ev1, c.stream) ;
        GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
        float ms = 0 ;
        cudaEventElapsedTime (&ms, c.ev0, c.ev1) ;
        R->info.device_ms = ms ; R->info.kernel_ms = ms ;
        return GB200_SUCCESS ;
    } ;
    gb200_status st = body () ;
    if (st != GB200_SUCCESS) { cudaStreamSynchronize (c.stream) ; cudaGetLastError () ; delete R ; return st ; }
    *out = R ;
    return GB200_SUCCESS ;
}

gb200_status gb200_select_host (gb200_result *out, const gb200_matrix *A, int select_op, int64_t k)
{
    if (out == NULL || A == NULL) return GB200_INVALID ;
    *out = NULL ;
    if (A->type_code < GB200_BOOL || A->type_code > GB200_FP64)
    {
        set_error ("operand of a user-defined type") ;
        return GB200_NOT_SUPPORTED ;
    }
    gb200_dmatrix dA = NULL ;
    bool cached = false ;
    GB200_TRY (cache_acquire (&dA, A, &cached)) ;
    gb200_status st = gb200_select_device (out, dA, select_op, k) ;
    if (cached) cache_release (dA) ; else gb200_dmatrix_free (&dA) ;
    return st ;
}

#pragma GCC visibility pop
} // extern "C"
