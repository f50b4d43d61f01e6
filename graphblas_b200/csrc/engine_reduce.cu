// engine_reduce.cu -- GrB_reduce of a matrix to a scalar on the device, and the hand-over of a
// multiply's result to the operand residency cache ("device-resident object cache +
// GrB_reduce to scalar": the call that follows the triangle-counting multiply,
// reference Demo/Source/tricount.c:177 -> Source/GB_reduce_to_scalar.c:107-270).
//
// The reduction itself is reduce_kernel (kernels_vec.cuh) over the monoid functors of semiring.cuh.
// Integer, boolean, MIN and MAX monoids are order-independent, so the scalar is the reference's bit for
// bit; floating-point PLUS / TIMES are summed in a fixed tree instead of the reference's left-to-right
// loop and agree within the reference's own 64 eps criterion.
#include "engine.cuh"
#include "kernels.cuh"

using namespace gb200 ;

extern "C" {
#pragma GCC visibility push(default)

gb200_status gb200_reduce_device (gb200_dmatrix Ad, int add_opcode, void *scalar)
{
    if (Ad == NULL || scalar == NULL) return GB200_INVALID ;
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    const DMat &A = Ad->v ;
    const int zc = A.type_code ;
    // the monoid must be one of the built-in ones for this type (boolean renames as in
    // gb200_semiring_canonical: MIN/TIMES -> LAND, MAX/PLUS -> LOR for bool)
    gb200_semiring s ;
    s.add_opcode = add_opcode ; s.mult_opcode = GB200_FIRST ; s.xy_code = zc ; s.z_code = zc ; s.flipxy = 0 ;
    GB200_TRY (gb200_semiring_canonical (&s)) ;
    int acc_size = 0 ;
    const uint64_t ident = identity_bits (zc, s.add_opcode, &acc_size) ;
    const int zsize = type_size (zc) ;
    uint64_t acc = ident ;
    if (A.nnz > 0)
    {
        int64_t grid = (A.nnz + 256 * 8 - 1) / (256 * 8) ;
        const int64_t cap = (int64_t) c.sm_count * 8 ;
        if (grid > cap) grid = cap ;
        if (grid < 1) grid = 1 ;
        DevBuf partial, out ;
        GB200_TRY (partial.alloc ((size_t) grid * 8)) ;
        GB200_TRY (out.alloc (16)) ;
        GB200_CUDA (cudaMemsetAsync (out.ptr, 0, 16, c.stream)) ;
        ReduceArgs ra ;
        ra.x = A.x ; ra.n = A.nnz ; ra.partial = partial.ptr ; ra.out = out.ptr ;
        ra.ticket = (unsigned int *) ((char *) out.ptr + 8) ;
        c.kev_used = 0 ;
        if (!launch_typed (zc, FAM_REDUCE, zc, s.add_opcode, GB200_FIRST, &ra, (int) grid, 256))
        { set_error ("no reduction for this monoid and type") ; return GB200_NOT_SUPPORTED ; }
        GB200_CUDA (cudaMemcpyAsync (c.pinned, out.ptr, 8, cudaMemcpyDeviceToHost, c.stream)) ;
        GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
        GB200_CUDA (cudaGetLastError ()) ;
        acc = 0 ;
        memcpy (&acc, c.pinned, (size_t) acc_size) ;
    }
    // accumulator -> the monoid's type (narrow types are accumulated in 32 bits and truncated)
    if (zc == GB200_BOOL) { const uint8_t b = ((uint32_t) acc != 0) ? 1 : 0 ; memcpy (scalar, &b, 1) ; }
    else memcpy (scalar, &acc, (size_t) zsize) ;            // little endian: the low bytes
    return GB200_SUCCESS ;
}

// the same over the values of a result that is still on the device: the checksum of a slab of C that is
// computed, summed and discarded (SURVEY.md 7 "hard parts": C = A*A on RMAT 24 exceeds HBM)
gb200_status gb200_result_reduce (gb200_result r, int add_opcode, void *scalar)
{
    if (r == NULL || scalar == NULL) return GB200_INVALID ;
    gb200_dmatrix_s view ;                      // borrows the result's value array
    view.v.x = r->x.ptr ; view.v.nnz = r->info.nnz ; view.v.type_code = r->info.type_code ;
    view.v.p = nullptr ; view.v.h = nullptr ; view.v.i = nullptr ;
    view.v.vlen = r->info.vlen ; view.v.vdim = r->info.vdim ; view.v.nvec = r->info.nvec ;
    view.v.hyper = 0 ; view.v.iso = 0 ; view.is_hyper_flag = 0 ;
    return gb200_reduce_device (&view, add_opcode, scalar) ;
}

gb200_status gb200_reduce_host (const gb200_matrix *A, int add_opcode, void *scalar)
{
    if (A == NULL || scalar == NULL) return GB200_INVALID ;
    if (A->type_code < GB200_BOOL || A->type_code > GB200_FP64)
    {
        set_error ("operand of a user-defined type") ;
        return GB200_NOT_SUPPORTED ;
    }
    gb200_dmatrix dA = NULL ;
    bool cached = false ;
    GB200_TRY (cache_acquire (&dA, A, &cached)) ;
    gb200_status st = gb200_reduce_device (dA, add_opcode, scalar) ;
    if (cached) cache_release (dA) ; else gb200_dmatrix_free (&dA) ;
    return st ;
}

// The caller has fetched T into the host arrays of `host` (gb200_result_fetch) and is done with the
// result handle: instead of freeing the device copy, remember it as the resident copy of those host
// arrays (if the residency cache is on and T is worth keeping), so that the next call on the same object
// -- GrB_reduce after the triangle-counting multiply, the next multiply of a k-truss loop -- starts
// from HBM.  Always consumes *r.
gb200_status gb200_result_adopt (gb200_result *r, const gb200_matrix *host)
{
    if (r == NULL || *r == NULL) return GB200_SUCCESS ;
    gb200_result_s *R = *r ;
    *r = NULL ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    const gb200_result_info &f = R->info ;
    const bool shape_ok = (host != NULL && host->vlen == f.vlen && host->vdim == f.vdim
        && host->nvec == f.nvec && host->type_code == f.type_code && (host->h != NULL) == (f.is_hyper != 0)) ;
    // the masked-dot kernels may read up to 32 bytes past the last index (aligned chunk loads)
    const bool slack_ok = (R->i.cap >= (size_t) (f.nnz + 8) * sizeof (int32_t)) ;
    if (shape_ok && slack_ok)
    {
        gb200_dmatrix_s *d = new (std::nothrow) gb200_dmatrix_s () ;
        if (d != NULL)
        {
            d->p = std::move (R->p) ; d->h = std::move (R->h) ; d->i = std::move (R->i) ; d->x = std::move (R->x) ;
            d->is_hyper_flag = f.is_hyper ;
            d->v.p = d->p.as<int64_t> () ;
            d->v.h = f.is_hyper ? d->h.as<int64_t> () : nullptr ;
            d->v.i = d->i.as<int32_t> () ;
            d->v.x = d->x.ptr ;
            d->v.vlen = f.vlen ; d->v.vdim = f.vdim ; d->v.nvec = f.nvec ; d->v.nnz = f.nnz ;
            d->v.hyper = (f.is_hyper && f.nvec < f.vdim) ? 1 : 0 ;
            d->v.type_code = f.type_code ;
            d->v.iso = 0 ;
            if (!cache_insert (d, host)) { gb200_dmatrix dd = d ; gb200_dmatrix_free (&dd) ; }
        }
    }
    delete R ;
    return GB200_SUCCESS ;
}

#pragma GCC visibility pop
} // extern "C"
