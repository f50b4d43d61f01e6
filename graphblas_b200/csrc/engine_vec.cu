// engine_vec.cu -- the vector multiplies (B is n-by-1): what GrB_mxv / GrB_vxm reach the seam with.
//
//   reference                                        | here
//   -------------------------------------------------+--------------------------------------------
//   GB_AxB_dot with B->vdim == 1                     | run_dotv: u expanded to values + presence
//     dot_nomask.c:20-83 / dot_compmask.c:20-126 /   |   bitmap, mask to a bitmap, dotv_kernel (one
//     dot_mask.c:33-159, inner dot_cij.c:102-148     |   lane group per vector of A) + dotv_long_kernel
//   GB_AxB_Gustavson / GB_AxB_heap with B->vdim == 1 | run_saxpyv: dense accumulator + presence bitmap,
//     Gustavson_nomask.c:66-159, heap_mask.c:190-426 |   saxpyv_kernel (one warp per entry of u) +
//                                                    |   saxpyv_long_kernel, then bitmap -> sorted list
//   GB_AxB_sequential.c:76-95 (mask policy)          | non-complemented mask: same rule; complemented
//                                                    |   mask: applied inside the kernel (the reference
//                                                    |   drops it and filters afterwards: same final C,
//                                                    |   SURVEY.md Appendix B.3), unless
//                                                    |   GB200_FUSE_COMPMASK=0
#include "engine.cuh"
#include "scan.cuh"
#include "kernels.cuh"

namespace gb200 {

static inline int grid_cap (int64_t n, int per_sm)
{
    int64_t cap = (int64_t) ctx ().sm_count * per_sm ;
    if (n > cap) n = cap ;
    if (n < 1) n = 1 ;
    return (int) n ;
}

// ---------------------------------------------------------------------------------------------
// long vectors of A, cut into segments (depends on A only: cached on the handle)
// ---------------------------------------------------------------------------------------------
__global__ void vec_nseg_kernel (const int64_t *__restrict__ p, int64_t nvec, int64_t *__restrict__ nseg)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < nvec ;
        t += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t len = p [t+1] - p [t] ;
        nseg [t] = (len > VEC_LONG) ? (len + VEC_LONG - 1) / VEC_LONG : 0 ;
    }
}

__global__ void vec_items_kernel (const int64_t *__restrict__ p, const int64_t *__restrict__ off,
    int64_t nvec, VecItem *__restrict__ items)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < nvec ;
        t += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t p0 = p [t], p1 = p [t+1] ;
        if (p1 - p0 <= VEC_LONG) continue ;
        int64_t q = off [t] ;
        for (int64_t s = p0 ; s < p1 ; s += VEC_LONG, q++)
        {
            VecItem it ;
            it.v = t ; it.p0 = s ; it.p1 = (s + VEC_LONG < p1) ? (s + VEC_LONG) : p1 ;
            items [q] = it ;
        }
    }
}

static gb200_status ensure_longitems (gb200_dmatrix_s *d)
{
    if (d->has_longitems) return GB200_SUCCESS ;
    Ctx &c = ctx () ;
    const int64_t nvec = d->v.nvec ;
    d->n_longitems = 0 ;
    if (nvec > 0)
    {
        DevBuf nseg, off ;
        GB200_TRY (nseg.alloc (nvec * sizeof (int64_t))) ;
        GB200_TRY (off.alloc ((nvec + 1) * sizeof (int64_t))) ;
        vec_nseg_kernel <<<grid_cap ((nvec + 255) / 256, 8), 256, 0, c.stream>>> (d->v.p, nvec,
            nseg.as<int64_t> ()) ;
        count_launch () ;
        GB200_TRY (scan_i64 (nseg.as<int64_t> (), off.as<int64_t> (), nvec)) ;
        int64_t total = 0 ;
        GB200_TRY (read_i64 (off.as<int64_t> () + nvec, &total)) ;
        if (total > 0)
        {
            GB200_TRY (d->longitems.alloc (total * sizeof (VecItem))) ;
            vec_items_kernel <<<grid_cap ((nvec + 255) / 256, 8), 256, 0, c.stream>>> (d->v.p,
                off.as<int64_t> (), nvec, d->longitems.as<VecItem> ()) ;
            count_launch () ;
        }
        GB200_CUDA (cudaGetLastError ()) ;
        d->n_longitems = total ;
    }
    d->has_longitems = true ;
    return GB200_SUCCESS ;
}

// tile_row [t] = the stored vector kk with p [kk] <= t * SPMV_TILE < p [kk+1]
__global__ void tile_row_kernel (const int64_t *__restrict__ p, int64_t nvec, int64_t ntiles,
    int32_t *__restrict__ tile_row)
{
    // entry ntiles: the vector that holds the last entry (trailing empty vectors are never visited)
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t <= ntiles ;
        t += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t pos = (t < ntiles) ? t * SPMV_TILE : (p [nvec] - 1) ;
        int64_t lo = 0, hi = nvec ;             // last kk in [0,nvec) with p [kk] <= pos
        while (hi - lo > 1)
        {
            const int64_t mid = (lo + hi) >> 1 ;
            if (__ldg (p + mid) <= pos) lo = mid ; else hi = mid ;
        }
        tile_row [t] = (int32_t) lo ;
    }
}

static gb200_status ensure_tilerow (gb200_dmatrix_s *d)
{
    if (d->has_tilerow) return GB200_SUCCESS ;
    Ctx &c = ctx () ;
    const int64_t ntiles = (d->v.nnz + SPMV_TILE - 1) / SPMV_TILE ;
    d->n_tiles = ntiles ;
    if (ntiles > 0)
    {
        GB200_TRY (d->tilerow.alloc ((ntiles + 1) * sizeof (int32_t))) ;
        tile_row_kernel <<<grid_cap ((ntiles + 256) / 256, 8), 256, 0, c.stream>>> (d->v.p, d->v.nvec,
            ntiles, d->tilerow.as<int32_t> ()) ;
        count_launch () ;
        GB200_CUDA (cudaGetLastError ()) ;
    }
    d->has_tilerow = true ;
    return GB200_SUCCESS ;
}

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
__global__ void bits_from_list_kernel (const int32_t *__restrict__ idx, int64_t n,
    uint32_t *__restrict__ bits)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < n ;
        t += (int64_t) gridDim.x * blockDim.x)
    {
        const uint32_t i = (uint32_t) idx [t] ;
        atomicOr (bits + (i >> 5), 1u << (i & 31)) ;
    }
}

// u (sparse) -> dense values + presence bitmap
template <class W>
__global__ void expand_vec_kernel_t (const int32_t *__restrict__ idx, const W *__restrict__ x,
    int64_t n, W *__restrict__ val, uint32_t *__restrict__ pres)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < n ;
        t += (int64_t) gridDim.x * blockDim.x)
    {
        const uint32_t i = (uint32_t) idx [t] ;
        val [i] = x [t] ;
        atomicOr (pres + (i >> 5), 1u << (i & 31)) ;
    }
}

// pack the flagged vectors of A into the result vector (ascending: stored order is name order)
__global__ void dotv_gather_kernel (const uint8_t *__restrict__ flags, const int64_t *__restrict__ pos,
    int64_t n, DMat A, const void *__restrict__ acc, int acc_size, int zsize, int is_bool,
    int32_t *__restrict__ Ci, void *__restrict__ Cx, int64_t *__restrict__ cum)
{
    for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; e <= n ;
        e += (int64_t) gridDim.x * blockDim.x)
    {
        if (e == n) { cum [0] = 0 ; cum [1] = pos [n] ; continue ; }
        if (!flags [e]) continue ;
        const int64_t q = pos [e] ;
        Ci [q] = (int32_t) dm_vecname (A, e) ;
        if (acc_size == 8) ((uint64_t *) Cx) [q] = ((const uint64_t *) acc) [e] ;
        else
        {
            const uint32_t a = ((const uint32_t *) acc) [e] ;
            if (is_bool) ((uint8_t *) Cx) [q] = (a != 0) ? 1 : 0 ;
            else if (zsize == 1) ((uint8_t *) Cx) [q] = (uint8_t) a ;
            else if (zsize == 2) ((uint16_t *) Cx) [q] = (uint16_t) a ;
            else ((uint32_t *) Cx) [q] = a ;
        }
    }
}

__global__ void pres_count_kernel (const uint32_t *__restrict__ pres, int64_t nwords,
    uint8_t *__restrict__ cnt)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < nwords ;
        t += (int64_t) gridDim.x * blockDim.x) cnt [t] = (uint8_t) __popc (pres [t]) ;
}

// presence bitmap + dense accumulator -> sorted index list + values
__global__ void pres_emit_kernel (const uint32_t *__restrict__ pres, const int64_t *__restrict__ pos,
    int64_t nwords, const void *__restrict__ acc, int acc_size, int zsize, int is_bool,
    int32_t *__restrict__ Ci, void *__restrict__ Cx, int64_t *__restrict__ cum)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t <= nwords ;
        t += (int64_t) gridDim.x * blockDim.x)
    {
        if (t == nwords) { cum [0] = 0 ; cum [1] = pos [nwords] ; continue ; }
        uint32_t w = pres [t] ;
        int64_t q = pos [t] ;
        while (w)
        {
            const int b = __ffs (w) - 1 ;
            w &= w - 1 ;
            const int64_t i = t * 32 + b ;
            Ci [q] = (int32_t) i ;
            if (acc_size == 8) ((uint64_t *) Cx) [q] = ((const uint64_t *) acc) [i] ;
            else
            {
                const uint32_t a = ((const uint32_t *) acc) [i] ;
                if (is_bool) ((uint8_t *) Cx) [q] = (a != 0) ? 1 : 0 ;
                else if (zsize == 1) ((uint8_t *) Cx) [q] = (uint8_t) a ;
                else if (zsize == 2) ((uint16_t *) Cx) [q] = (uint16_t) a ;
                else ((uint32_t *) Cx) [q] = a ;
            }
            q++ ;
        }
    }
}

// structural mask vector -> bitmap over [0,n)
static gb200_status mask_bitmap (const gb200_dmatrix_s *M, int64_t n, DevBuf &mbits)
{
    Ctx &c = ctx () ;
    DMat Mv ; DevBuf Mp2, Mi2 ;
    GB200_TRY (filter_mask (M, Mv, Mp2, Mi2)) ;
    const int64_t nwords = (n + 31) / 32 ;
    GB200_TRY (mbits.alloc ((nwords > 0 ? nwords : 1) * sizeof (uint32_t))) ;
    GB200_CUDA (cudaMemsetAsync (mbits.ptr, 0, mbits.bytes, c.stream)) ;
    if (Mv.nnz > 0)
    {
        bits_from_list_kernel <<<grid_cap ((Mv.nnz + 255) / 256, 8), 256, 0, c.stream>>> (Mv.i, Mv.nnz,
            mbits.as<uint32_t> ()) ;
        count_launch () ;
        GB200_CUDA (cudaGetLastError ()) ;
    }
    // Mp2 / Mi2 are freed in stream order after the kernel above
    return GB200_SUCCESS ;
}

bool vec_shape (const gb200_dmatrix_s *A, const gb200_dmatrix_s *B)
{
    return B->v.vdim == 1 && B->v.nvec == 1 && !B->v.hyper && B->v.nnz > 0 && A->v.nvec > 0
        && A->v.nnz > 0 ;
}

// =============================================================================================
// pull: w<M> = A'*u
// =============================================================================================
gb200_status run_dotv (gb200_result_s *R, const gb200_dmatrix_s *M, int mask_comp,
    const gb200_dmatrix_s *Ad, const gb200_dmatrix_s *Bd, const gb200_semiring &s)
{
    Ctx &c = ctx () ;
    // no mask: every vector of A is computed, so the entries of A are streamed in equal tiles;
    // with a mask whole vectors are skipped, which the vector-per-lane-group kernel does for free
    const char *senv = getenv ("GB200_SPMV_STREAM") ;
    const bool stream = (M == nullptr) && !(senv != nullptr && atoi (senv) == 0) ;
    if (stream) GB200_TRY (ensure_tilerow (const_cast<gb200_dmatrix_s *> (Ad))) ;
    else GB200_TRY (ensure_longitems (const_cast<gb200_dmatrix_s *> (Ad))) ;
    const DMat &A = Ad->v ;
    const DMat &B = Bd->v ;
    const int64_t cvlen = A.vdim, vlen = A.vlen, anvec = A.nvec ;
    R->info.method_used = GB200_METHOD_DOT ;
    R->info.type_code = s.z_code ;
    R->info.mask_applied = (M != nullptr) ? 1 : 0 ;             // GB_AxB_dot.c:315
    int acc_size = 0 ;
    (void) identity_bits (s.z_code, s.add_opcode, &acc_size) ;
    const int zsize = type_size (s.z_code) ;
    const int tsz = type_size (B.type_code) ;

    // ---- u: in place if every entry is present, else expanded --------------------------------
    DevBuf bval, bpres, mbits ;
    const void *bv = B.x ;
    const uint32_t *bp = nullptr ;
    if (B.nnz != vlen)
    {
        const int64_t nwords = (vlen + 31) / 32 ;
        GB200_TRY (bval.alloc ((size_t) vlen * tsz)) ;
        GB200_TRY (bpres.alloc (nwords * sizeof (uint32_t))) ;
        GB200_CUDA (cudaMemsetAsync (bpres.ptr, 0, bpres.bytes, c.stream)) ;
        const int g = grid_cap ((B.nnz + 255) / 256, 8) ;
        switch (tsz)
        {
            case 1 : expand_vec_kernel_t<uint8_t>  <<<g, 256, 0, c.stream>>> (B.i, (const uint8_t *)  B.x, B.nnz, bval.as<uint8_t> (),  bpres.as<uint32_t> ()) ; break ;
            case 2 : expand_vec_kernel_t<uint16_t> <<<g, 256, 0, c.stream>>> (B.i, (const uint16_t *) B.x, B.nnz, bval.as<uint16_t> (), bpres.as<uint32_t> ()) ; break ;
            case 4 : expand_vec_kernel_t<uint32_t> <<<g, 256, 0, c.stream>>> (B.i, (const uint32_t *) B.x, B.nnz, bval.as<uint32_t> (), bpres.as<uint32_t> ()) ; break ;
            default: expand_vec_kernel_t<uint64_t> <<<g, 256, 0, c.stream>>> (B.i, (const uint64_t *) B.x, B.nnz, bval.as<uint64_t> (), bpres.as<uint32_t> ()) ; break ;
        }
        count_launch () ;
        bv = bval.ptr ; bp = bpres.as<uint32_t> () ;
    }
    if (M != nullptr) GB200_TRY (mask_bitmap (M, cvlen, mbits)) ;

    DevBuf vals, flags, pos, nmatch, Ci, Cx, ccum ;
    GB200_TRY (vals.alloc ((size_t) anvec * acc_size)) ;
    GB200_TRY (flags.alloc (anvec)) ;
    GB200_TRY (pos.alloc ((anvec + 1) * sizeof (int64_t))) ;
    GB200_TRY (nmatch.alloc (8)) ;
    GB200_CUDA (cudaMemsetAsync (nmatch.ptr, 0, 8, c.stream)) ;

    if (stream)
    {
        int asz = 0 ;
        const uint64_t ident = identity_bits (s.z_code, s.add_opcode, &asz) ;
        GB200_TRY (fill_bits (vals.ptr, acc_size, ident, anvec)) ;
        GB200_CUDA (cudaMemsetAsync (flags.ptr, 0, anvec, c.stream)) ;
        SpmvArgs sp ;
        memset (&sp, 0, sizeof (sp)) ;
        sp.A = A ; sp.bval = bv ; sp.bpres = bp ;
        sp.tile_row = Ad->tilerow.as<int32_t> () ; sp.ntiles = Ad->n_tiles ;
        sp.vals = vals.ptr ; sp.flags = flags.as<uint8_t> () ;
        sp.nmatch = nmatch.as<unsigned long long> () ;
        sp.mult_op = s.mult_opcode ; sp.flip = s.flipxy ;
        const char *oenv = getenv ("GB200_SPMV_OCC8") ;            // 1: the 32-register build, 8 blocks per SM
        const char *genv = getenv ("GB200_SPMV_GRID") ;            // blocks per SM
        const bool occ8 = (oenv != nullptr && atoi (oenv) != 0 && bp == nullptr) ;
        const int per_sm = (genv != nullptr && atoi (genv) > 0) ? atoi (genv) : (occ8 ? 8 : 6) ;
        if (!launch_typed (s.xy_code, bp ? FAM_SPMV_PRES : (occ8 ? FAM_SPMV_OCC8 : FAM_SPMV), s.z_code,
            s.add_opcode, s.mult_opcode, &sp, grid_cap (sp.ntiles, per_sm), SPMV_THREADS))
        { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
    }
    DotVArgs da ;
    memset (&da, 0, sizeof (da)) ;
    da.A = A ; da.bval = bv ; da.bpres = bp ;
    da.mbits = (M != nullptr) ? mbits.as<uint32_t> () : nullptr ;
    da.mask_comp = mask_comp ;
    da.vals = vals.ptr ; da.flags = flags.as<uint8_t> () ;
    da.items = Ad->longitems.as<VecItem> () ; da.nitems = Ad->n_longitems ;
    da.nmatch = nmatch.as<unsigned long long> () ;
    da.mult_op = s.mult_opcode ; da.flip = s.flipxy ;
    const double avg = (double) A.nnz / (double) anvec ;
    int G = 4 ; while (G < 32 && 4 * G <= avg) G <<= 1 ;
    if (getenv ("GB200_DOTV_G")) G = atoi (getenv ("GB200_DOTV_G")) ;
    da.G = G ;
    const int64_t gpb = 256 / G ;
    if (!stream && !launch_typed (s.xy_code, FAM_DOTV, s.z_code, s.add_opcode, s.mult_opcode, &da,
        grid_cap ((anvec + gpb - 1) / gpb, 8), 256))
    { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
    if (!stream && da.nitems > 0)
    {
        if (!launch_typed (s.xy_code, FAM_DOTV_LONG, s.z_code, s.add_opcode, s.mult_opcode, &da,
            grid_cap ((da.nitems + 7) / 8, 8), 256))
        { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
    }
    GB200_CUDA (cudaGetLastError ()) ;

    // ---- flags -> scan -> pack ----------------------------------------------------------------
    GB200_TRY (scan_u8 (flags.as<uint8_t> (), pos.as<int64_t> (), anvec)) ;
    int64_t cnz = 0 ;
    GB200_TRY (read_i64 (pos.as<int64_t> () + anvec, &cnz)) ;
    GB200_TRY (Ci.alloc ((cnz > 0 ? cnz : 1) * sizeof (int32_t))) ;
    GB200_TRY (Cx.alloc ((size_t) (cnz > 0 ? cnz : 1) * zsize)) ;
    GB200_TRY (ccum.alloc (2 * sizeof (int64_t))) ;
    dotv_gather_kernel <<<grid_cap ((anvec + 256) / 256, 8), 256, 0, c.stream>>> (flags.as<uint8_t> (),
        pos.as<int64_t> (), anvec, A, vals.ptr, acc_size, zsize, s.z_code == GB200_BOOL,
        Ci.as<int32_t> (), Cx.ptr, ccum.as<int64_t> ()) ;
    count_launch () ;
    GB200_CUDA (cudaGetLastError ()) ;
    GB200_TRY (assemble (R, 1, nullptr, false, ccum, Ci, Cx, cnz, false, cvlen, 1)) ;
    int64_t nm = 0 ;
    GB200_TRY (read_i64 (nmatch.as<int64_t> (), &nm)) ;
    R->info.flops = nm ;
    return GB200_SUCCESS ;
}

// =============================================================================================
// push: w<M> = A*u
// =============================================================================================
gb200_status run_saxpyv (gb200_result_s *R, const gb200_dmatrix_s *Min, int mask_comp,
    const gb200_dmatrix_s *Ad, const gb200_dmatrix_s *Bd, const gb200_semiring &s)
{
    Ctx &c = ctx () ;
    const DMat &A = Ad->v ;
    const DMat &B = Bd->v ;
    const int64_t cvlen = A.vlen ;
    // one GPU saxpy serves GUSTAVSON and HEAP requests alike; an explicit HEAP request is reported as
    // HEAP, as GB_AxB_select.c:139-143 would (its automatic heap choice, :95-128, is a CPU workspace
    // trade-off that has no analogue here and is reported as GUSTAVSON: INTEGRATION.md)
    R->info.method_used = (ctx ().method_request == GB200_METHOD_HEAP) ? GB200_METHOD_HEAP : GB200_METHOD_GUSTAVSON ;
    R->info.type_code = s.z_code ;

    // ---- mask policy ----------------------------------------------------------------------------
    const gb200_dmatrix_s *M = Min ;
    int64_t masked_flops = -1 ;
    if (M != nullptr && mask_comp)
    {
        // the reference's saxpy ignores a complemented mask (GB_AxB_sequential.c:76-81) and GB_mxm
        // filters T afterwards; applying it here gives the same C without materialising the
        // entries that would be deleted
        const char *env = getenv ("GB200_FUSE_COMPMASK") ;
        if (env != nullptr && atoi (env) == 0) M = nullptr ;
    }
    else if (M != nullptr && c.mask_policy == 2) M = nullptr ;
    else if (M != nullptr)
    {
        DevBuf flops, cum ;
        GB200_TRY (flopcount (&M->v, A, B, flops, cum, &masked_flops)) ;
        if (masked_flops <= M->v.nnz && c.mask_policy != 1) M = nullptr ;   // GB_AxB_sequential.c:88-95
    }
    R->info.mask_applied = (M != nullptr) ? 1 : 0 ;

    int acc_size = 0 ;
    const uint64_t ident = identity_bits (s.z_code, s.add_opcode, &acc_size) ;
    const int zsize = type_size (s.z_code) ;
    const int64_t nwords = (cvlen + 31) / 32 ;

    DevBuf acc, pres, mbits, longlist, counters, cntw, pos, Ci, Cx, ccum ;
    GB200_TRY (acc.alloc ((size_t) cvlen * acc_size)) ;
    GB200_TRY (fill_bits (acc.ptr, acc_size, ident, cvlen)) ;
    GB200_TRY (pres.alloc (nwords * sizeof (uint32_t))) ;
    GB200_CUDA (cudaMemsetAsync (pres.ptr, 0, pres.bytes, c.stream)) ;
    if (M != nullptr) GB200_TRY (mask_bitmap (M, cvlen, mbits)) ;
    GB200_TRY (longlist.alloc (B.nnz * sizeof (int32_t))) ;
    GB200_TRY (counters.alloc (16)) ;
    GB200_CUDA (cudaMemsetAsync (counters.ptr, 0, 16, c.stream)) ;

    SaxpyVArgs sa ;
    memset (&sa, 0, sizeof (sa)) ;
    sa.A = A ; sa.B = B ;
    sa.mbits = (M != nullptr) ? mbits.as<uint32_t> () : nullptr ;
    sa.mask_comp = mask_comp ;
    sa.acc = acc.ptr ; sa.pres = pres.as<uint32_t> () ;
    sa.longlist = longlist.as<int32_t> () ;
    sa.hugelist = longlist.as<int32_t> () + (B.nnz - 1) ;
    sa.nlong = (unsigned int *) counters.ptr ;
    sa.nflops = (unsigned long long *) ((char *) counters.ptr + 8) ;
    sa.mult_op = s.mult_opcode ; sa.flip = s.flipxy ;
    if (!launch_typed (s.xy_code, FAM_SAXPYV, s.z_code, s.add_opcode, s.mult_opcode, &sa,
        grid_cap ((B.nnz + 7) / 8, 8), 256))
    { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
    if (!launch_typed (s.xy_code, FAM_SAXPYV_LONG, s.z_code, s.add_opcode, s.mult_opcode, &sa,
        c.sm_count * 4, 256))
    { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
    GB200_CUDA (cudaGetLastError ()) ;

    // ---- bitmap -> sorted list ------------------------------------------------------------------
    GB200_TRY (cntw.alloc (nwords)) ;
    GB200_TRY (pos.alloc ((nwords + 1) * sizeof (int64_t))) ;
    pres_count_kernel <<<grid_cap ((nwords + 255) / 256, 8), 256, 0, c.stream>>> (pres.as<uint32_t> (),
        nwords, cntw.as<uint8_t> ()) ;
    count_launch () ;
    GB200_TRY (scan_u8 (cntw.as<uint8_t> (), pos.as<int64_t> (), nwords)) ;
    int64_t cnz = 0 ;
    GB200_TRY (read_i64 (pos.as<int64_t> () + nwords, &cnz)) ;
    GB200_TRY (Ci.alloc ((cnz > 0 ? cnz : 1) * sizeof (int32_t))) ;
    GB200_TRY (Cx.alloc ((size_t) (cnz > 0 ? cnz : 1) * zsize)) ;
    GB200_TRY (ccum.alloc (2 * sizeof (int64_t))) ;
    pres_emit_kernel <<<grid_cap ((nwords + 256) / 256, 8), 256, 0, c.stream>>> (pres.as<uint32_t> (),
        pos.as<int64_t> (), nwords, acc.ptr, acc_size, zsize, s.z_code == GB200_BOOL,
        Ci.as<int32_t> (), Cx.ptr, ccum.as<int64_t> ()) ;
    count_launch () ;
    GB200_CUDA (cudaGetLastError ()) ;
    int64_t nf = 0 ;
    GB200_TRY (read_i64 ((const int64_t *) sa.nflops, &nf)) ;
    R->info.flops = (masked_flops >= 0 && M != nullptr) ? masked_flops : nf ;
    return assemble (R, 1, nullptr, false, ccum, Ci, Cx, cnz, false, cvlen, 1) ;
}

} // namespace gb200
