// engine_accum.cu -- C<M> = accum (C,T) on the device (the step immediately AFTER the
// multiply for every masked / accumulated call, reference Source/GB_accum_mask.c:130-328 -> GB_add.c
// (Z = accum (C,T)) -> GB_mask.c (C<M> = Z), MATLAB statements Test/GB_spec_accum.m, GB_spec_mask.m:60-90).
//
// Reference behaviour restated, entry by entry, for C, T and M held in the same orientation:
//   Z = T cast to C's type                                  when there is no accumulator, else
//   Z(i,j) = accum (C(i,j) cast to x, T(i,j) cast to y), cast from z to C's type, where both exist;
//            C(i,j) where only C has the entry; T(i,j) cast to C's type where only T has it (GB_add.c);
//   m(i,j) = M(i,j) exists and its value cast to bool is true, negated for a complemented mask; without a
//            mask m is true everywhere (GB_mask.c:181-190);
//   result: where m, the entry of Z (or no entry if Z has none); elsewhere the entry of C, or no entry when
//            C_replace (GB_spec_mask.m:60-90).
// The reference's other route (GB_subassign_kernel when T is the smaller operand, GB_accum_mask.c:230-248)
// ends in the same matrix once its zombies and pending tuples are assembled.
//
// GPU: nothing is merged sequentially.  Every entry of C looks its position up in T (binary search in the
// vector of the same name) and in M; every entry of T looks itself up in C, and those that C lacks look up M.
// Each entry then knows whether it survives and where its value comes from.  Two scans over the survivors
// (one over C's entries, one over T's entries that C lacks) give the output position of every survivor in
// closed form: position = survivors of C before me + survivors of T-only before my insertion point, because
// both arrays list their vectors in ascending order.  The values are cast whole-array by the library's
// cast kernels (the GB_CAST rule) and combined by the multiply operators of semiring.cuh.
#include "engine.cuh"
#include "scan.cuh"
#include "semiring.cuh"

namespace gb200 {

enum { AM_FROM_C = 0, AM_FROM_T = 1, AM_BOTH = 2 } ;

static inline int am_grid (int64_t n, int per_sm = 16)
{
    int64_t g = (n + 255) / 256, cap = (int64_t) ctx ().sm_count * per_sm ;
    if (g > cap) g = cap ;
    if (g < 1) g = 1 ;
    return (int) g ;
}

// entries [pa, pe) of the vector named j; an absent vector gives the empty range at the place it would
// occupy (so that prefix sums taken at pa count exactly the entries of the vectors before j)
__device__ __forceinline__ void am_range (const DMat &X, int64_t j, int64_t &pa, int64_t &pe)
{
    if (!X.hyper) { pa = __ldg (X.p + j) ; pe = __ldg (X.p + j + 1) ; return ; }
    int64_t lo = 0, hi = X.nvec ;
    while (lo < hi)
    {
        const int64_t mid = (lo + hi) >> 1 ;
        if (__ldg (X.h + mid) < j) lo = mid + 1 ; else hi = mid ;
    }
    pa = __ldg (X.p + lo) ;
    pe = (lo < X.nvec && __ldg (X.h + lo) == j) ? __ldg (X.p + lo + 1) : pa ;
}

// first position in idx [lo, hi) whose index is >= key
__device__ __forceinline__ int64_t am_lower (const int32_t *__restrict__ idx, int64_t lo, int64_t hi, int32_t key)
{
    while (lo < hi)
    {
        const int64_t mid = (lo + hi) >> 1 ;
        if (__ldg (idx + mid) < key) lo = mid + 1 ; else hi = mid ;
    }
    return lo ;
}

__device__ __forceinline__ bool am_mask (const DMat &M, int has_mask, int mask_comp, int64_t j, int32_t i)
{
    if (!has_mask) return !mask_comp ;
    int64_t m0, m1 ;
    am_range (M, j, m0, m1) ;
    const int64_t u = am_lower (M.i, m0, m1, i) ;
    const bool present = (u < m1 && __ldg (M.i + u) == i) ;      // false-valued entries were filtered out
    return present != (mask_comp != 0) ;
}

// an entry of C: does it survive, and as what
__global__ void am_c_kernel (DMat C, DMat T, DMat M, const int32_t *__restrict__ vecC, int has_mask,
    int mask_comp, int replace, int has_accum, uint8_t *__restrict__ keepC, uint8_t *__restrict__ kindC,
    int64_t *__restrict__ posT)
{
    for (int64_t pc = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; pc < C.nnz ;
        pc += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t j = vecC [pc] ;
        const int32_t i = __ldg (C.i + pc) ;
        int64_t t0, t1 ;
        am_range (T, j, t0, t1) ;
        const int64_t u = am_lower (T.i, t0, t1, i) ;
        const bool tex = (u < t1 && __ldg (T.i + u) == i) ;
        const bool m = am_mask (M, has_mask, mask_comp, j, i) ;
        bool keep ;
        int kind ;
        if (m)
        {
            // the entry of Z
            keep = has_accum ? true : tex ;
            kind = has_accum ? (tex ? AM_BOTH : AM_FROM_C) : AM_FROM_T ;
        }
        else
        {
            keep = !replace ;
            kind = AM_FROM_C ;
        }
        keepC [pc] = keep ? 1 : 0 ;
        kindC [pc] = (uint8_t) kind ;
        posT [pc] = u ;
    }
}

// an entry of T that C lacks survives where the mask admits it
__global__ void am_t_kernel (DMat C, DMat T, DMat M, const int32_t *__restrict__ vecT, int has_mask,
    int mask_comp, uint8_t *__restrict__ keepT, int64_t *__restrict__ posC)
{
    for (int64_t pt = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; pt < T.nnz ;
        pt += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t j = vecT [pt] ;
        const int32_t i = __ldg (T.i + pt) ;
        int64_t c0, c1 ;
        am_range (C, j, c0, c1) ;
        const int64_t v = am_lower (C.i, c0, c1, i) ;
        const bool cex = (v < c1 && __ldg (C.i + v) == i) ;
        keepT [pt] = (!cex && am_mask (M, has_mask, mask_comp, j, i)) ? 1 : 0 ;
        posC [pt] = v ;
    }
}

// survivors -> their place in R; srcC / srcT: where the value comes from (-1: not from there)
__global__ void am_fill_c_kernel (DMat C, const uint8_t *__restrict__ keepC, const uint8_t *__restrict__ kindC,
    const int64_t *__restrict__ posT, const int64_t *__restrict__ sC, const int64_t *__restrict__ sT,
    int32_t *__restrict__ Ri, int64_t *__restrict__ srcC, int64_t *__restrict__ srcT)
{
    for (int64_t pc = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; pc < C.nnz ;
        pc += (int64_t) gridDim.x * blockDim.x)
    {
        if (!keepC [pc]) continue ;
        const int64_t u = posT [pc], q = sC [pc] + sT [u] ;
        const int kind = kindC [pc] ;
        Ri [q] = __ldg (C.i + pc) ;
        srcC [q] = (kind == AM_FROM_T) ? -1 : pc ;
        srcT [q] = (kind == AM_FROM_C) ? -1 : u ;
    }
}

__global__ void am_fill_t_kernel (DMat T, const uint8_t *__restrict__ keepT, const int64_t *__restrict__ posC,
    const int64_t *__restrict__ sC, const int64_t *__restrict__ sT, int32_t *__restrict__ Ri,
    int64_t *__restrict__ srcC, int64_t *__restrict__ srcT)
{
    for (int64_t pt = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; pt < T.nnz ;
        pt += (int64_t) gridDim.x * blockDim.x)
    {
        if (!keepT [pt]) continue ;
        const int64_t q = sT [pt] + sC [posC [pt]] ;
        Ri [q] = __ldg (T.i + pt) ;
        srcC [q] = -1 ;
        srcT [q] = pt ;
    }
}

// cum [j] = survivors in the vectors before j, for every vector name j of R (cum [vdim] = all of them)
__global__ void am_cum_kernel (DMat C, DMat T, const int64_t *__restrict__ sC, const int64_t *__restrict__ sT,
    int64_t vdim, int64_t *__restrict__ cum)
{
    for (int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; j <= vdim ;
        j += (int64_t) gridDim.x * blockDim.x)
    {
        if (j == vdim) { cum [j] = sC [C.nnz] + sT [T.nnz] ; continue ; }
        int64_t c0, c1, t0, t1 ;
        am_range (C, j, c0, c1) ;
        am_range (T, j, t0, t1) ;
        cum [j] = sC [c0] + sT [t0] ;
    }
}

// z [q] = op (x [srcC [q]], y [srcT [q]]) where both sources exist
template <class X, class Z>
__global__ void am_op_kernel (const int64_t *__restrict__ srcC, const int64_t *__restrict__ srcT, int64_t n,
    const X *__restrict__ x, const X *__restrict__ y, int op, Z *__restrict__ z)
{
    for (int64_t q = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; q < n ; q += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t a = srcC [q], b = srcT [q] ;
        if (a >= 0 && b >= 0) z [q] = mult_one<X, Z> (op, x [a], y [b], false) ;
        else z [q] = Z (0) ;
    }
}

template <class X>
static void am_op_launch (const int64_t *srcC, const int64_t *srcT, int64_t n, const void *x, const void *y,
    int op, bool z_is_bool, void *z, cudaStream_t st)
{
    const int g = am_grid (n) ;
    if (z_is_bool && !std::is_same<X, bool>::value)
        am_op_kernel<X, bool> <<<g, 256, 0, st>>> (srcC, srcT, n, (const X *) x, (const X *) y, op, (bool *) z) ;
    else
        am_op_kernel<X, X> <<<g, 256, 0, st>>> (srcC, srcT, n, (const X *) x, (const X *) y, op, (X *) z) ;
}

// Rx [q] = the value of survivor q: C's own, T's cast to C's type, or the accumulated one (all tsz bytes)
__global__ void am_value_kernel (const int64_t *__restrict__ srcC, const int64_t *__restrict__ srcT, int64_t n,
    const unsigned char *__restrict__ Cx, const unsigned char *__restrict__ Tc,
    const unsigned char *__restrict__ Zc, int tsz, unsigned char *__restrict__ Rx)
{
    for (int64_t q = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; q < n ; q += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t a = srcC [q], b = srcT [q] ;
        const unsigned char *src = (a >= 0 && b >= 0) ? (Zc + q * tsz) : ((a >= 0) ? (Cx + a * tsz) : (Tc + b * tsz)) ;
        if (tsz == 8) ((uint64_t *) Rx) [q] = *(const uint64_t *) src ;
        else if (tsz == 4) ((uint32_t *) Rx) [q] = *(const uint32_t *) src ;
        else if (tsz == 2) ((uint16_t *) Rx) [q] = *(const uint16_t *) src ;
        else Rx [q] = *src ;
    }
}

static int am_boolean_rename (int op)           // Source/GB_boolean_rename.c:30-91
{
    switch (op)
    {
        case GB200_DIV : case GB200_FIRST : return GB200_FIRST ;
        case GB200_MIN : case GB200_TIMES : case GB200_LAND : return GB200_LAND ;
        case GB200_MAX : case GB200_PLUS : case GB200_LOR : return GB200_LOR ;
        case GB200_MINUS : case GB200_ISNE : case GB200_NE : case GB200_LXOR : return GB200_LXOR ;
        case GB200_ISEQ : case GB200_EQ : return GB200_EQ ;
        case GB200_ISGT : case GB200_GT : return GB200_GT ;
        case GB200_ISLT : case GB200_LT : return GB200_LT ;
        case GB200_ISGE : case GB200_GE : return GB200_GE ;
        case GB200_ISLE : case GB200_LE : return GB200_LE ;
        default : return op ;
    }
}

} // namespace gb200

using namespace gb200 ;

extern "C" {
#pragma GCC visibility push(default)

gb200_status gb200_accum_mask_device (gb200_result *out, gb200_dmatrix Cd, gb200_dmatrix Td, gb200_dmatrix Md,
    int mask_comp, int c_replace, int accum_opcode, int accum_xy_code, int result_hyper)
{
    if (out == NULL || Cd == NULL || Td == NULL) return GB200_INVALID ;
    *out = NULL ;
    const DMat &C = Cd->v ;
    const DMat &T = Td->v ;
    if (C.vlen != T.vlen || C.vdim != T.vdim || (Md != NULL && (Md->v.vlen != C.vlen || Md->v.vdim != C.vdim)))
        return GB200_INVALID ;
    const int has_accum = (accum_opcode != 0) ;
    if (has_accum && (accum_opcode < GB200_FIRST || accum_opcode > GB200_LE || accum_xy_code < GB200_BOOL
        || accum_xy_code > GB200_FP64))
    {
        set_error ("accumulator outside the built-in operator/type space") ;
        return GB200_NOT_SUPPORTED ;
    }
    if (C.vdim > ((int64_t) 1 << 27))
    {
        set_error ("accum/mask over %lld vectors: the vector space is walked densely", (long long) C.vdim) ;
        return GB200_NOT_SUPPORTED ;
    }
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    gb200_result_s *R = new (std::nothrow) gb200_result_s () ;
    if (R == NULL) return GB200_OUT_OF_MEMORY ;
    memset (&R->info, 0, sizeof (R->info)) ;
    auto body = [&] () -> gb200_status
    {
        cudaEventRecord (c.ev0, c.stream) ;
        const int64_t nC = C.nnz, nT = T.nnz ;
        const int ctype = C.type_code, csz = type_size (ctype) ;
        // the mask as a structure: entries whose value is false do not admit (GB_mask.c cast_M)
        DMat M ;
        memset (&M, 0, sizeof (M)) ;
        DevBuf Mp2, Mi2 ;
        if (Md != NULL) GB200_TRY (filter_mask (Md, M, Mp2, Mi2)) ;
        DevBuf vecC, vecT, keepC, kindC, keepT, posT, posC, sC, sT ;
        GB200_TRY (vecC.alloc ((size_t) (nC > 0 ? nC : 1) * sizeof (int32_t))) ;
        GB200_TRY (vecT.alloc ((size_t) (nT > 0 ? nT : 1) * sizeof (int32_t))) ;
        GB200_TRY (keepC.alloc (nC > 0 ? nC : 1)) ;
        GB200_TRY (kindC.alloc (nC > 0 ? nC : 1)) ;
        GB200_TRY (keepT.alloc (nT > 0 ? nT : 1)) ;
        GB200_TRY (posT.alloc ((size_t) (nC > 0 ? nC : 1) * sizeof (int64_t))) ;
        GB200_TRY (posC.alloc ((size_t) (nT > 0 ? nT : 1) * sizeof (int64_t))) ;
        GB200_TRY (sC.alloc ((size_t) (nC + 1) * sizeof (int64_t))) ;
        GB200_TRY (sT.alloc ((size_t) (nT + 1) * sizeof (int64_t))) ;
        GB200_TRY (launch_vecof (C, vecC.as<int32_t> ())) ;
        GB200_TRY (launch_vecof (T, vecT.as<int32_t> ())) ;
        if (nC > 0)
        {
            am_c_kernel <<<am_grid (nC), 256, 0, c.stream>>> (C, T, M, vecC.as<int32_t> (), Md != NULL, mask_comp,
                c_replace, has_accum, keepC.as<uint8_t> (), kindC.as<uint8_t> (), posT.as<int64_t> ()) ;
            count_launch () ;
        }
        if (nT > 0)
        {
            am_t_kernel <<<am_grid (nT), 256, 0, c.stream>>> (C, T, M, vecT.as<int32_t> (), Md != NULL, mask_comp,
                keepT.as<uint8_t> (), posC.as<int64_t> ()) ;
            count_launch () ;
        }
        GB200_CUDA (cudaGetLastError ()) ;
        GB200_TRY (scan_u8 (keepC.as<uint8_t> (), sC.as<int64_t> (), nC)) ;
        GB200_TRY (scan_u8 (keepT.as<uint8_t> (), sT.as<int64_t> (), nT)) ;
        int64_t kc = 0, kt = 0 ;
        GB200_TRY (read_i64 (sC.as<int64_t> () + nC, &kc)) ;
        GB200_TRY (read_i64 (sT.as<int64_t> () + nT, &kt)) ;
        const int64_t rnz = kc + kt ;
        DevBuf Ri, srcC, srcT, cum, Rx ;
        GB200_TRY (Ri.alloc ((size_t) (rnz > 0 ? rnz : 1) * sizeof (int32_t))) ;
        GB200_TRY (srcC.alloc ((size_t) (rnz > 0 ? rnz : 1) * sizeof (int64_t))) ;
        GB200_TRY (srcT.alloc ((size_t) (rnz > 0 ? rnz : 1) * sizeof (int64_t))) ;
        GB200_TRY (cum.alloc ((size_t) (C.vdim + 1) * sizeof (int64_t))) ;
        GB200_TRY (Rx.alloc ((size_t) (rnz > 0 ? rnz : 1) * csz)) ;
        if (nC > 0)
        {
            am_fill_c_kernel <<<am_grid (nC), 256, 0, c.stream>>> (C, keepC.as<uint8_t> (), kindC.as<uint8_t> (),
                posT.as<int64_t> (), sC.as<int64_t> (), sT.as<int64_t> (), Ri.as<int32_t> (),
                srcC.as<int64_t> (), srcT.as<int64_t> ()) ;
            count_launch () ;
        }
        if (nT > 0)
        {
            am_fill_t_kernel <<<am_grid (nT), 256, 0, c.stream>>> (T, keepT.as<uint8_t> (), posC.as<int64_t> (),
                sC.as<int64_t> (), sT.as<int64_t> (), Ri.as<int32_t> (), srcC.as<int64_t> (),
                srcT.as<int64_t> ()) ;
            count_launch () ;
        }
        am_cum_kernel <<<am_grid (C.vdim + 1, 8), 256, 0, c.stream>>> (C, T, sC.as<int64_t> (), sT.as<int64_t> (),
            C.vdim, cum.as<int64_t> ()) ;
        count_launch () ;
        GB200_CUDA (cudaGetLastError ()) ;
        // values
        if (rnz > 0)
        {
            DevBuf Tc, Cxx, Tyy, Zb, Zc ;
            const void *tc = T.x ;
            if (nT > 0 && T.type_code != ctype)
            {
                GB200_TRY (cast_values (T.x, T.type_code, ctype, nT, Tc)) ;
                tc = Tc.ptr ;
            }
            const void *zc = nullptr ;
            if (has_accum)
            {
                int op = accum_opcode ;
                const int xy = accum_xy_code ;
                if (xy == GB200_BOOL) op = am_boolean_rename (op) ;
                const bool z_is_bool = (xy == GB200_BOOL) || (op >= GB200_EQ) ;
                const int zcode = z_is_bool ? GB200_BOOL : xy ;
                const void *cx = C.x, *ty = T.x ;
                if (nC > 0 && ctype != xy) { GB200_TRY (cast_values (C.x, ctype, xy, nC, Cxx)) ; cx = Cxx.ptr ; }
                if (nT > 0 && T.type_code != xy) { GB200_TRY (cast_values (T.x, T.type_code, xy, nT, Tyy)) ; ty = Tyy.ptr ; }
                GB200_TRY (Zb.alloc ((size_t) rnz * type_size (zcode))) ;
                const int64_t *sc = srcC.as<int64_t> (), *stt = srcT.as<int64_t> () ;
                switch (xy)
                {
                    case GB200_BOOL   : am_op_launch<bool>     (sc, stt, rnz, cx, ty, op, z_is_bool, Zb.ptr, c.stream) ; break ;
                    case GB200_INT8   : am_op_launch<int8_t>   (sc, stt, rnz, cx, ty, op, z_is_bool, Zb.ptr, c.stream) ; break ;
                    case GB200_UINT8  : am_op_launch<uint8_t>  (sc, stt, rnz, cx, ty, op, z_is_bool, Zb.ptr, c.stream) ; break ;
                    case GB200_INT16  : am_op_launch<int16_t>  (sc, stt, rnz, cx, ty, op, z_is_bool, Zb.ptr, c.stream) ; break ;
                    case GB200_UINT16 : am_op_launch<uint16_t> (sc, stt, rnz, cx, ty, op, z_is_bool, Zb.ptr, c.stream) ; break ;
                    case GB200_INT32  : am_op_launch<int32_t>  (sc, stt, rnz, cx, ty, op, z_is_bool, Zb.ptr, c.stream) ; break ;
                    case GB200_UINT32 : am_op_launch<uint32_t> (sc, stt, rnz, cx, ty, op, z_is_bool, Zb.ptr, c.stream) ; break ;
                    case GB200_INT64  : am_op_launch<int64_t>  (sc, stt, rnz, cx, ty, op, z_is_bool, Zb.ptr, c.stream) ; break ;
                    case GB200_UINT64 : am_op_launch<uint64_t> (sc, stt, rnz, cx, ty, op, z_is_bool, Zb.ptr, c.stream) ; break ;
                    case GB200_FP32   : am_op_launch<float>    (sc, stt, rnz, cx, ty, op, z_is_bool, Zb.ptr, c.stream) ; break ;
                    default           : am_op_launch<double>   (sc, stt, rnz, cx, ty, op, z_is_bool, Zb.ptr, c.stream) ; break ;
                }
                count_launch () ;
                GB200_CUDA (cudaGetLastError ()) ;
                zc = Zb.ptr ;
                if (zcode != ctype) { GB200_TRY (cast_values (Zb.ptr, zcode, ctype, rnz, Zc)) ; zc = Zc.ptr ; }
            }
            am_value_kernel <<<am_grid (rnz), 256, 0, c.stream>>> (srcC.as<int64_t> (), srcT.as<int64_t> (), rnz,
                (const unsigned char *) C.x, (const unsigned char *) tc, (const unsigned char *) zc, csz,
                (unsigned char *) Rx.ptr) ;
            count_launch () ;
            GB200_CUDA (cudaGetLastError ()) ;
            GB200_CUDA (cudaStreamSynchronize (c.stream)) ;         // the cast buffers die at scope exit
        }
        R->info.type_code = ctype ;
        R->info.method_used = 0 ; R->info.mask_applied = (Md != NULL) ; R->info.flops = nC + nT ;
        GB200_TRY (assemble (R, C.vdim, nullptr, false, cum, Ri, Rx, rnz, result_hyper != 0, C.vlen, C.vdim)) ;
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

gb200_status gb200_accum_mask_host (gb200_result *out, const gb200_matrix *C, const gb200_matrix *T,
    const gb200_matrix *M, int mask_comp, int c_replace, int accum_opcode, int accum_xy_code, int result_hyper)
{
    if (out == NULL || C == NULL || T == NULL) return GB200_INVALID ;
    *out = NULL ;
    const gb200_matrix *ops [3] = { C, T, M } ;
    for (int k = 0 ; k < 3 ; k++)
        if (ops [k] != NULL && (ops [k]->type_code < GB200_BOOL || ops [k]->type_code > GB200_FP64))
        {
            set_error ("operand of a user-defined type") ;
            return GB200_NOT_SUPPORTED ;
        }
    gb200_dmatrix d [3] = { NULL, NULL, NULL } ;
    bool cached [3] = { false, false, false } ;
    gb200_status st = GB200_SUCCESS ;
    for (int k = 0 ; k < 3 && st == GB200_SUCCESS ; k++)
    {
        if (ops [k] == NULL) continue ;
        if (k == 2 && M == C) { d [2] = d [0] ; continue ; }       // C<C> = ...: one copy
        st = cache_acquire (&d [k], ops [k], &cached [k]) ;
    }
    if (st == GB200_SUCCESS)
        st = gb200_accum_mask_device (out, d [0], d [1], d [2], mask_comp, c_replace, accum_opcode,
            accum_xy_code, result_hyper) ;
    for (int k = 0 ; k < 3 ; k++)
    {
        if (d [k] == NULL || (k == 2 && M == C)) continue ;
        if (cached [k]) cache_release (d [k]) ; else gb200_dmatrix_free (&d [k]) ;
    }
    return st ;
}

gb200_status gb200_assign_scalar_device (gb200_result *out, gb200_dmatrix Cd, gb200_dmatrix Md, int c_replace,
    int accum_opcode, int accum_xy_code, const void *scalar, int scalar_code, int result_hyper)
{
    if (out == NULL || Cd == NULL || Md == NULL || scalar == NULL) return GB200_INVALID ;
    *out = NULL ;
    if (scalar_code < GB200_BOOL || scalar_code > GB200_FP64)
    {
        set_error ("scalar of a user-defined type") ;
        return GB200_NOT_SUPPORTED ;
    }
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    // T = the scalar on the pattern of the mask's true entries: the expanded scalar is dense, but C<M> = Z only
    // ever looks at it where the mask admits
    DMat Mv ;
    DevBuf Mp2, Mi2, Tx ;
    GB200_TRY (filter_mask (Md, Mv, Mp2, Mi2)) ;
    const int tsz = type_size (scalar_code) ;
    GB200_TRY (Tx.alloc ((size_t) (Mv.nnz > 0 ? Mv.nnz : 1) * tsz)) ;
    uint64_t bits = 0 ;
    memcpy (&bits, scalar, tsz) ;
    GB200_TRY (fill_bits (Tx.ptr, tsz, bits, Mv.nnz)) ;
    gb200_dmatrix_s Tt ;
    Tt.v = Mv ;
    Tt.v.x = Tx.ptr ;
    Tt.v.type_code = scalar_code ;
    Tt.v.iso = 1 ;
    Tt.is_hyper_flag = Md->is_hyper_flag ;
    Tt.iso_known = 1 ;
    return gb200_accum_mask_device (out, Cd, &Tt, Md, 0, c_replace, accum_opcode, accum_xy_code, result_hyper) ;
}

gb200_status gb200_assign_scalar_host (gb200_result *out, const gb200_matrix *C, const gb200_matrix *M,
    int c_replace, int accum_opcode, int accum_xy_code, const void *scalar, int scalar_code, int result_hyper)
{
    if (out == NULL || C == NULL || M == NULL || scalar == NULL) return GB200_INVALID ;
    *out = NULL ;
    if (C->type_code < GB200_BOOL || C->type_code > GB200_FP64 || M->type_code < GB200_BOOL
        || M->type_code > GB200_FP64)
    {
        set_error ("operand of a user-defined type") ;
        return GB200_NOT_SUPPORTED ;
    }
    gb200_dmatrix dC = NULL, dM = NULL ;
    bool cached_c = false, cached_m = false ;
    GB200_TRY (cache_acquire (&dC, C, &cached_c)) ;
    gb200_status st = cache_acquire (&dM, M, &cached_m) ;
    if (st == GB200_SUCCESS)
        st = gb200_assign_scalar_device (out, dC, dM, c_replace, accum_opcode, accum_xy_code, scalar, scalar_code,
            result_hyper) ;
    if (dM != NULL) { if (cached_m) cache_release (dM) ; else gb200_dmatrix_free (&dM) ; }
    if (cached_c) cache_release (dC) ; else gb200_dmatrix_free (&dC) ;
    return st ;
}

#pragma GCC visibility pop
} // extern "C"
