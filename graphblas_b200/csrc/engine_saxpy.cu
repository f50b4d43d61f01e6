// engine_saxpy.cu -- C=A*B and C<M>=A*B by the saxpy formulation: the GPU replacement of
// GB_AxB_flopcount, GB_AxB_Gustavson (symbolic + numeric) and GB_AxB_heap.
//
//   reference                                   | here
//   --------------------------------------------+------------------------------------------------
//   GB_AxB_flopcount.c:167-295                  | flop_light_kernel / flop_long_kernel + scan
//   GB_AxB_select.c bjnz_max scan :70-76        | not needed (no heap/Gustavson choice on the GPU)
//   Gustavson_symbolic.c:44-256 (+GB_qsort_1)   | sym_hash_kernel (shared-memory hash set + bitonic
//                                               | sort, binned by flops) / bitmap for heavy vectors
//   Gustavson_nomask.c:66-159, _mask.c:93-285   | saxpy_light_kernel / saxpy_heavy_kernel (kernels.cuh)
//   GB_AxB_sequential.c:76-95 (mask policy)     | run_saxpy: complemented mask dropped, mask dropped
//                                               | when flops <= nnz(M)
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

// =============================================================================================
// flop count
// =============================================================================================
constexpr int64_t FLOP_LONG = 4096 ;        // vectors of B longer than this use the multi-block kernel

__device__ __forceinline__ int64_t entry_flops (const DMat &A, int64_t k, bool masked,
    int64_t im_first, int64_t im_last)
{
    int64_t pa, pe ;
    if (!dm_lookup (A, k, pa, pe)) return 0 ;
    if (masked)
    {
        // skip A(:,k) when its index range cannot meet the mask's (GB_AxB_flopcount.c:262-270)
        const int64_t alo = __ldg (A.i + pa), ahi = __ldg (A.i + pe - 1) ;
        if (ahi < im_first || alo > im_last) return 0 ;
    }
    return pe - pa ;
}

__device__ __forceinline__ bool mask_range (const DMat &M, int64_t j, int64_t &im_first, int64_t &im_last)
{
    int64_t pm, pme ;
    if (!dm_lookup (M, j, pm, pme)) return false ;
    im_first = __ldg (M.i + pm) ; im_last = __ldg (M.i + pme - 1) ;
    return true ;
}

// G lanes per stored vector of B
__global__ void flop_light_kernel (DMat A, DMat B, DMat M, int masked, int G,
    int64_t *__restrict__ flops, int32_t *__restrict__ longlist, unsigned int *__restrict__ nlong)
{
    const int gl = threadIdx.x & (G - 1) ;
    const int64_t gpb = blockDim.x / G ;
    const int64_t stride = (int64_t) gridDim.x * gpb ;
    const int64_t nvec = B.nvec ;
    const int64_t niter = (nvec + stride - 1) / stride ;
    int64_t kk = (int64_t) blockIdx.x * gpb + threadIdx.x / G ;
    for (int64_t itn = 0 ; itn < niter ; itn++, kk += stride)
    {
        int64_t f = 0 ;
        bool is_long = false ;
        if (kk < nvec)
        {
            const int64_t pb0 = B.p [kk], pb1 = B.p [kk+1] ;
            if (pb1 - pb0 > FLOP_LONG) is_long = true ;
            else if (pb1 > pb0)
            {
                int64_t im_first = 0, im_last = 0 ;
                bool go = true ;
                if (masked) go = mask_range (M, dm_vecname (B, kk), im_first, im_last) ;
                if (go)
                    for (int64_t pb = pb0 + gl ; pb < pb1 ; pb += G)
                        f += entry_flops (A, B.i [pb], masked, im_first, im_last) ;
            }
        }
        for (int off = G >> 1 ; off > 0 ; off >>= 1) f += __shfl_down_sync (0xffffffffu, f, off, G) ;
        if (kk < nvec && gl == 0)
        {
            if (is_long) { flops [kk] = 0 ; longlist [atomicAdd (nlong, 1u)] = (int32_t) kk ; }
            else flops [kk] = f ;
        }
    }
}

// grid (x, nlong): block x of long vector y strides over its entries
__global__ void flop_long_kernel (DMat A, DMat B, DMat M, int masked,
    const int32_t *__restrict__ longlist, int64_t *__restrict__ flops)
{
    __shared__ int64_t ws [33] ;
    const int64_t kk = longlist [blockIdx.y] ;
    const int64_t pb0 = B.p [kk], pb1 = B.p [kk+1] ;
    int64_t im_first = 0, im_last = 0 ;
    bool go = true ;
    if (masked) go = mask_range (M, dm_vecname (B, kk), im_first, im_last) ;
    int64_t f = 0 ;
    if (go)
        for (int64_t pb = pb0 + blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; pb < pb1 ;
            pb += (int64_t) gridDim.x * blockDim.x)
            f += entry_flops (A, B.i [pb], masked, im_first, im_last) ;
    int64_t total ;
    block_excl_scan_i64 (f, ws, total) ;
    if (threadIdx.x == 0 && total) atomicAdd ((unsigned long long *) (flops + kk), (unsigned long long) total) ;
}

// flops: per stored vector of B (size nvec); cum: its exclusive scan (size nvec+1)
gb200_status flopcount (const DMat *M, const DMat &A, const DMat &B, DevBuf &flops, DevBuf &cum,
    int64_t *total)
{
    Ctx &c = ctx () ;
    const int64_t nvec = B.nvec ;
    GB200_TRY (flops.alloc ((nvec > 0 ? nvec : 1) * sizeof (int64_t))) ;
    GB200_TRY (cum.alloc ((nvec + 1) * sizeof (int64_t))) ;
    if (nvec > 0)
    {
        DevBuf longlist, nlong ;
        GB200_TRY (longlist.alloc (nvec * sizeof (int32_t))) ;
        GB200_TRY (nlong.alloc (8)) ;
        GB200_CUDA (cudaMemsetAsync (nlong.ptr, 0, 8, c.stream)) ;
        const double avg = (double) B.nnz / (double) nvec ;
        int G = 1 ; while (G < 32 && G < avg) G <<= 1 ;
        DMat Mv = M ? *M : DMat () ;
        const int64_t groups = 256 / G ;
        flop_light_kernel <<<grid_cap ((nvec + groups - 1) / groups, 16), 256, 0, c.stream>>> (A, B, Mv,
            M != nullptr, G, flops.as<int64_t> (), longlist.as<int32_t> (), nlong.as<unsigned int> ()) ;
        count_launch () ;
        int64_t nl = 0 ;
        GB200_TRY (read_i64 (nlong.as<int64_t> (), &nl)) ;
        nl &= 0xffffffffLL ;
        if (nl > 0)
        {
            // y dimension of a grid is limited to 65535
            for (int64_t y0 = 0 ; y0 < nl ; y0 += 65535)
            {
                const int64_t ny = (nl - y0 < 65535) ? (nl - y0) : 65535 ;
                dim3 grid ((unsigned) ((nl <= 8) ? c.sm_count * 2 : 16), (unsigned) ny) ;
                flop_long_kernel <<<grid, 256, 0, c.stream>>> (A, B, Mv, M != nullptr,
                    longlist.as<int32_t> () + y0, flops.as<int64_t> ()) ;
                count_launch () ;
            }
        }
        GB200_CUDA (cudaGetLastError ()) ;
    }
    GB200_TRY (scan_i64 (flops.as<int64_t> (), cum.as<int64_t> (), nvec)) ;
    GB200_TRY (read_i64 (cum.as<int64_t> () + nvec, total)) ;
    return GB200_SUCCESS ;
}

// =============================================================================================
// binning of B's vectors by flops
// =============================================================================================
constexpr int NCLASS = 5 ;                  // 4 shared-memory classes + heavy
constexpr int CL_HEAVY = 4 ;
static const int64_t class_limit [4] = { 128, 1024, 8192, 16384 } ;    // flops upper bounds
static const int class_log [4]       = { 8, 11, 14, 15 } ;             // hash table = 2 x limit
static const int class_threads [4]   = { 32, 128, 256, 512 } ;

struct ClassLimits { int64_t lim [4] ; } ;

// Position of the calling thread's vector in the list of class cl (cl < 0: none): the lanes of a warp
// that append to the same class share one atomic (a matrix whose vectors all fall in one class --
// Erdos-Renyi -- made every thread hit the same counter: 0.7 ms for 1 M vectors).  Called by whole warps.
__device__ __forceinline__ unsigned int class_append (unsigned int *counts, int cl, int nclass)
{
    const int lane = threadIdx.x & 31 ;
    unsigned int pos = 0 ;
    for (int q = 0 ; q < nclass ; q++)
    {
        const unsigned peers = __ballot_sync (0xffffffffu, cl == q) ;
        if (peers == 0) continue ;
        unsigned int base = 0 ;
        if (lane == __ffs (peers) - 1) base = atomicAdd (counts + q, (unsigned int) __popc (peers)) ;
        base = __shfl_sync (0xffffffffu, base, __ffs (peers) - 1) ;
        if (cl == q) pos = base + __popc (peers & ((1u << lane) - 1u)) ;
    }
    return pos ;
}

__global__ void classify_kernel (const int64_t *__restrict__ flops, int64_t nvec, ClassLimits L,
    int32_t *__restrict__ lists, unsigned int *__restrict__ counts)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x ;
    const int64_t niter = (nvec + stride - 1) / stride ;
    int64_t kk = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ;
    for (int64_t it = 0 ; it < niter ; it++, kk += stride)
    {
        int cl = -1 ;
        if (kk < nvec)
        {
            const int64_t f = flops [kk] ;
            if (f > 0)
            {
                cl = CL_HEAVY ;
                #pragma unroll
                for (int q = 3 ; q >= 0 ; q--) if (f <= L.lim [q]) cl = q ;
            }
        }
        const unsigned int pos = class_append (counts, cl, CL_HEAVY + 1) ;
        if (cl >= 0) lists [(int64_t) cl * nvec + pos] = (int32_t) kk ;
    }
}

// after the count pass: the vectors of the four shared-memory flop classes again, by their number of
// ENTRIES -- the fused fill + numeric kernel (saxpy_hash_kernel) keeps (row, accumulator) slots, so
// what has to fit its table is the pattern, not the flops.  Classes 3 and 4: the pattern is too long
// for it (flops <= 8 Ki / <= 16 Ki: the two-kernel path with the shared-memory set of that flop class).
struct CntLimits { int64_t lim [3] ; int64_t mid_flops ; int64_t heavy_flops ; } ;

__global__ void classify_cnt_kernel (const int64_t *__restrict__ flops, const int64_t *__restrict__ Cp,
    int64_t nvec, CntLimits L, int32_t *__restrict__ lists, unsigned int *__restrict__ counts)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x ;
    const int64_t niter = (nvec + stride - 1) / stride ;
    int64_t kk = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ;
    for (int64_t it = 0 ; it < niter ; it++, kk += stride)
    {
        int cl = -1 ;
        if (kk < nvec)
        {
            const int64_t f = flops [kk] ;
            const int64_t cnt = Cp [kk+1] - Cp [kk] ;
            if (f > 0 && f <= L.heavy_flops && cnt > 0)
            {
                cl = (f > L.mid_flops) ? 4 : 3 ;    // pattern too long: by flop class again
                #pragma unroll
                for (int q = 2 ; q >= 0 ; q--) if (cnt <= L.lim [q]) cl = q ;
            }
        }
        const unsigned int pos = class_append (counts, cl, 5) ;
        if (cl >= 0) lists [(int64_t) cl * nvec + pos] = (int32_t) kk ;
    }
}

// =============================================================================================
// symbolic phase, shared-memory classes: hash set of row indices, then (FILL) compaction + bitonic
// sort so that the indices of C(:,j) come out ascending (the invariant checked by the reference at
// Source/GB_matvec_check.c:521-523)
// =============================================================================================
template <bool FILL>
__global__ void sym_hash_kernel (DMat A, DMat B, const int32_t *__restrict__ cols, int64_t ncols,
    int LOG, int64_t *__restrict__ cnt, const int64_t *__restrict__ Cp, int32_t *__restrict__ Ci)
{
    extern __shared__ int32_t sm [] ;
    int32_t *table = sm ;
    const int size = 1 << LOG ;
    int32_t *comp = sm + size ;                 // FILL only: size/2 entries
    __shared__ int s_n ;
    const uint32_t mask = size - 1 ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5 ;
    for (int64_t c = blockIdx.x ; c < ncols ; c += gridDim.x)
    {
        const int64_t kk = cols [c] ;
        for (int t = threadIdx.x ; t < size ; t += blockDim.x) table [t] = -1 ;
        if (threadIdx.x == 0) s_n = 0 ;
        __syncthreads () ;
        int mine = 0 ;
        const int64_t pb0 = B.p [kk], pb1 = B.p [kk+1] ;
        for_each_product (A, B, pb0, pb1, [&] (int64_t p, int64_t)
        {
            const int32_t i = __ldg (A.i + p) ;
            uint32_t h = (hash32 ((uint32_t) i) >> (32 - LOG)) & mask ;
            while (true)
            {
                const int32_t old = atomicCAS (table + h, -1, i) ;
                if (old == -1) { mine++ ; break ; }
                if (old == i) break ;
                h = (h + 1) & mask ;
            }
        }) ;
        if (!FILL)
        {
            for (int off = 16 ; off > 0 ; off >>= 1) mine += __shfl_down_sync (0xffffffffu, mine, off) ;
            if (lane == 0 && mine) atomicAdd (&s_n, mine) ;
            __syncthreads () ;
            if (threadIdx.x == 0) cnt [kk] = s_n ;
            __syncthreads () ;
        }
        else
        {
            __syncthreads () ;
            for (int t = threadIdx.x ; t < size ; t += blockDim.x)
            {
                const int32_t key = table [t] ;
                if (key >= 0) comp [atomicAdd (&s_n, 1)] = key ;
            }
            __syncthreads () ;
            const int n = s_n ;
            int n2 = 1 ; while (n2 < n) n2 <<= 1 ;
            for (int t = n + threadIdx.x ; t < n2 ; t += blockDim.x) comp [t] = INT32_MAX ;
            __syncthreads () ;
            for (int k = 2 ; k <= n2 ; k <<= 1)
            {
                for (int j = k >> 1 ; j > 0 ; j >>= 1)
                {
                    for (int t = threadIdx.x ; t < n2 ; t += blockDim.x)
                    {
                        const int ixj = t ^ j ;
                        if (ixj > t)
                        {
                            const int32_t a = comp [t], b = comp [ixj] ;
                            const bool asc = ((t & k) == 0) ;
                            if ((a > b) == asc) { comp [t] = b ; comp [ixj] = a ; }
                        }
                    }
                    __syncthreads () ;
                }
            }
            const int64_t base = Cp [kk] ;
            for (int t = threadIdx.x ; t < n ; t += blockDim.x) Ci [base + t] = comp [t] ;
            __syncthreads () ;
        }
    }
}

// =============================================================================================
// symbolic phase when the index range is small (vlen <= SYMB_MAX_VLEN): a vlen-bit bitmap in shared
// memory instead of a hash set.  Marking is one shared atomicOr per product; the pattern comes out of
// the bitmap already ascending, so the FILL pass needs no sort.
// =============================================================================================
constexpr int64_t SYMB_MAX_VLEN = 512 * 1024 ;          // 64 KB of shared memory

template <bool FILL>
__global__ void sym_bitmap_kernel (DMat A, DMat B, const int32_t *__restrict__ cols, int64_t ncols,
    int nwords, int64_t *__restrict__ cnt, const int64_t *__restrict__ Cp, int32_t *__restrict__ Ci)
{
    extern __shared__ uint32_t sbm [] ;
    __shared__ int64_t ws [33] ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5 ;
    for (int64_t c = blockIdx.x ; c < ncols ; c += gridDim.x)
    {
        const int64_t kk = cols [c] ;
        for (int t = threadIdx.x ; t < nwords ; t += blockDim.x) sbm [t] = 0u ;
        __syncthreads () ;
        const int64_t pb0 = B.p [kk], pb1 = B.p [kk+1] ;
        for_each_product (A, B, pb0, pb1, [&] (int64_t p, int64_t)
        {
            const uint32_t i = (uint32_t) __ldg (A.i + p) ;
            const uint32_t bit = 1u << (i & 31) ;
            if (!(sbm [i >> 5] & bit)) atomicOr (sbm + (i >> 5), bit) ;
        }) ;
        __syncthreads () ;
        if (!FILL)
        {
            int64_t mine = 0 ;
            for (int t = threadIdx.x ; t < nwords ; t += blockDim.x) mine += __popc (sbm [t]) ;
            int64_t total ;
            block_excl_scan_i64 (mine, ws, total) ;
            if (threadIdx.x == 0) cnt [kk] = total ;
        }
        else
        {
            int64_t run = Cp [kk] ;
            for (int t0 = 0 ; t0 < nwords ; t0 += blockDim.x)
            {
                const int t = t0 + threadIdx.x ;
                uint32_t word = (t < nwords) ? sbm [t] : 0u ;
                int64_t total ;
                int64_t q = run + block_excl_scan_i64 (__popc (word), ws, total) ;
                while (word)
                {
                    const int b = __ffs (word) - 1 ;
                    Ci [q++] = (int32_t) (t * 32 + b) ;
                    word &= word - 1 ;
                }
                run += total ;
            }
        }
        __syncthreads () ;
    }
}

// =============================================================================================
// heavy vectors: one bitmap of vlen bits per vector in flight
// =============================================================================================
// split the heavy vectors of one batch into work items of about `target` flops each
__global__ void heavy_items_kernel (DMat B, const int32_t *__restrict__ heavy, int64_t nh,
    const int64_t *__restrict__ flops, int64_t target, int64_t wbase, HeavyItem *__restrict__ items,
    unsigned int *__restrict__ nitems)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < nh ;
        t += (int64_t) gridDim.x * blockDim.x)
    {
        const int32_t kk = heavy [t] ;
        const int64_t pb0 = B.p [kk], pb1 = B.p [kk+1], bjnz = pb1 - pb0 ;
        int64_t nchunks = (flops [kk] + target - 1) / target ;
        if (nchunks > bjnz) nchunks = bjnz ;
        if (nchunks < 1) nchunks = 1 ;
        const int64_t ch = (bjnz + nchunks - 1) / nchunks ;
        nchunks = (bjnz + ch - 1) / ch ;
        const unsigned int base = atomicAdd (nitems, (unsigned int) nchunks) ;
        for (int64_t q = 0 ; q < nchunks ; q++)
        {
            HeavyItem it ;
            it.kk = kk ; it.w = (int32_t) (wbase + t) ;
            it.pb0 = pb0 + q * ch ;
            it.pb1 = (pb0 + (q + 1) * ch < pb1) ? (pb0 + (q + 1) * ch) : pb1 ;
            items [base + q] = it ;
        }
    }
}

// the same split with the items laid out in the order of the heavy list (item range of vector t =
// [ioff [t], ioff [t+1])), so that a run of vectors is a run of items: nch == nullptr fills the items
__device__ __forceinline__ int64_t heavy_chunking (const DMat &B, int32_t kk, int64_t f, int64_t target,
    int64_t &ch)
{
    const int64_t bjnz = B.p [kk+1] - B.p [kk] ;
    int64_t nchunks = (f + target - 1) / target ;
    if (nchunks > bjnz) nchunks = bjnz ;
    if (nchunks < 1) nchunks = 1 ;
    ch = (bjnz + nchunks - 1) / nchunks ;
    return (bjnz + ch - 1) / ch ;
}

__global__ void heavy_items_ordered_kernel (DMat B, const int32_t *__restrict__ heavy, int64_t nh,
    const int64_t *__restrict__ flops, int64_t target, int64_t *__restrict__ nch,
    const int64_t *__restrict__ ioff, HeavyItem *__restrict__ items)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < nh ;
        t += (int64_t) gridDim.x * blockDim.x)
    {
        const int32_t kk = heavy [t] ;
        int64_t ch ;
        const int64_t nchunks = heavy_chunking (B, kk, flops [kk], target, ch) ;
        if (nch != nullptr) { nch [t] = nchunks ; continue ; }
        const int64_t pb0 = B.p [kk], pb1 = B.p [kk+1] ;
        for (int64_t q = 0 ; q < nchunks ; q++)
        {
            HeavyItem it ;
            it.kk = kk ; it.w = (int32_t) t ;
            it.pb0 = pb0 + q * ch ;
            it.pb1 = (pb0 + (q + 1) * ch < pb1) ? (pb0 + (q + 1) * ch) : pb1 ;
            items [ioff [t] + q] = it ;
        }
    }
}

// set the bit of every row index reached by the products of the item
__global__ void heavy_mark_kernel (DMat A, DMat B, const HeavyItem *__restrict__ items,
    int64_t nitems, uint32_t *__restrict__ bitmap, int64_t nwords)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5 ;
    for (int64_t it = blockIdx.x ; it < nitems ; it += gridDim.x)
    {
        const HeavyItem item = items [it] ;
        uint32_t *bm = bitmap + (int64_t) item.w * nwords ;
        for (int64_t pb = item.pb0 + warp ; pb < item.pb1 ; pb += nwarps)
        {
            int64_t pa, pe ;
            if (!dm_lookup (A, B.i [pb], pa, pe)) continue ;
            for (int64_t p = pa + lane ; p < pe ; p += 32)
            {
                const uint32_t i = (uint32_t) __ldg (A.i + p) ;
                const uint32_t bit = 1u << (i & 31) ;
                if (!(bm [i >> 5] & bit)) atomicOr (bm + (i >> 5), bit) ;
            }
        }
    }
}

// masked: the bitmap of a heavy vector is the pattern of M(:,j)
__global__ void heavy_mark_list_kernel (const int32_t *__restrict__ heavy, int64_t nh,
    const int64_t *__restrict__ lp, const int32_t *__restrict__ li, const int64_t *__restrict__ lpos,
    uint32_t *__restrict__ bitmap, int64_t nwords)
{
    for (int64_t w = blockIdx.x ; w < nh ; w += gridDim.x)
    {
        const int64_t kk = heavy [w] ;
        const int64_t lv = lpos ? lpos [kk] : kk ;
        if (lv < 0) continue ;
        uint32_t *bm = bitmap + w * nwords ;
        for (int64_t p = lp [lv] + threadIdx.x ; p < lp [lv+1] ; p += blockDim.x)
        {
            const uint32_t i = (uint32_t) li [p] ;
            atomicOr (bm + (i >> 5), 1u << (i & 31)) ;
        }
    }
}

// dst [t] = src [idx [t]]
__global__ void gather_i64_kernel (const int64_t *__restrict__ src, const int32_t *__restrict__ idx,
    int64_t n, int64_t *__restrict__ dst)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < n ;
        t += (int64_t) gridDim.x * blockDim.x) dst [t] = src [idx [t]] ;
}

// nnz of each heavy vector = popcount of its bitmap
__global__ void heavy_count_kernel (const int32_t *__restrict__ heavy, int64_t nh,
    const uint32_t *__restrict__ bitmap, int64_t nwords, int64_t *__restrict__ cnt)
{
    __shared__ int64_t ws [33] ;
    for (int64_t w = blockIdx.x ; w < nh ; w += gridDim.x)
    {
        const uint32_t *bm = bitmap + w * nwords ;
        int64_t s = 0 ;
        for (int64_t t = threadIdx.x ; t < nwords ; t += blockDim.x) s += __popc (bm [t]) ;
        int64_t total ;
        block_excl_scan_i64 (s, ws, total) ;
        if (threadIdx.x == 0) cnt [heavy [w]] = total ;
    }
}

// rank[word] = number of set bits in the words before it; optionally emit the (ascending) indices
__global__ void heavy_rank_kernel (const int32_t *__restrict__ heavy, int64_t nh,
    const uint32_t *__restrict__ bitmap, int32_t *__restrict__ rank, int64_t nwords,
    const int64_t *__restrict__ Cp, int32_t *__restrict__ Ci)
{
    __shared__ int64_t ws [33] ;
    __shared__ int64_t s_run ;
    for (int64_t w = blockIdx.x ; w < nh ; w += gridDim.x)
    {
        const uint32_t *bm = bitmap + w * nwords ;
        int32_t *rk = rank + w * nwords ;
        const int64_t base = Ci ? Cp [heavy [w]] : 0 ;
        if (threadIdx.x == 0) s_run = 0 ;
        __syncthreads () ;
        for (int64_t t0 = 0 ; t0 < nwords ; t0 += blockDim.x)
        {
            const int64_t t = t0 + threadIdx.x ;
            const uint32_t word = (t < nwords) ? bm [t] : 0u ;
            int64_t total ;
            const int64_t ex = block_excl_scan_i64 (__popc (word), ws, total) ;
            const int64_t r = s_run + ex ;
            if (t < nwords)
            {
                rk [t] = (int32_t) r ;
                if (Ci)
                {
                    uint32_t wv = word ; int64_t q = base + r ;
                    while (wv)
                    {
                        const int b = __ffs (wv) - 1 ;
                        Ci [q++] = (int32_t) (t * 32 + b) ;
                        wv &= wv - 1 ;
                    }
                }
            }
            __syncthreads () ;
            if (threadIdx.x == 0) s_run += total ;
            __syncthreads () ;
        }
    }
}

// position of vector j = name of B's kk-th vector in the mask's pointer array, or -1
__global__ void mask_pos_kernel (DMat B, DMat M, int64_t *__restrict__ lpos)
{
    for (int64_t kk = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; kk < B.nvec ;
        kk += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t j = dm_vecname (B, kk) ;
        int64_t r = -1 ;
        if (!M.hyper) r = j ;
        else
        {
            int64_t lo = 0, hi = M.nvec - 1 ;
            while (lo <= hi)
            {
                const int64_t mid = (lo + hi) >> 1, hv = M.h [mid] ;
                if (hv == j) { r = mid ; break ; }
                if (hv < j) lo = mid + 1 ; else hi = mid - 1 ;
            }
        }
        lpos [kk] = r ;
    }
}

gb200_status launch_mask_pos (const DMat &B, const DMat &M, int64_t *lpos)
{
    if (B.nvec <= 0) return GB200_SUCCESS ;
    mask_pos_kernel <<<grid_cap ((B.nvec + 255) / 256, 8), 256, 0, ctx ().stream>>> (B, M, lpos) ;
    count_launch () ;
    GB200_CUDA (cudaGetLastError ()) ;
    return GB200_SUCCESS ;
}

// masked saxpy: per stored vector of B, how many flagged slots does its mask vector hold
__global__ void masked_counts_kernel (const int64_t *__restrict__ lpos, const int64_t *__restrict__ lp,
    const int64_t *__restrict__ pos, int64_t nvec, int64_t *__restrict__ cnt)
{
    for (int64_t kk = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; kk < nvec ;
        kk += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t lv = lpos [kk] ;
        cnt [kk] = (lv < 0) ? 0 : (pos [lp [lv+1]] - pos [lp [lv]]) ;
    }
}

// gather flagged slots (mask order == ascending index order) into Ci / Cx
__global__ void masked_gather_kernel (const int64_t *__restrict__ lpos, const int64_t *__restrict__ lp,
    const int32_t *__restrict__ li, const uint8_t *__restrict__ flags, const int64_t *__restrict__ pos,
    const int64_t *__restrict__ cum, const void *__restrict__ acc, int acc_size, int zsize, int is_bool,
    int64_t nvec, int32_t *__restrict__ Ci, void *__restrict__ Cx)
{
    // one warp per stored vector of B
    const int lane = threadIdx.x & 31 ;
    const int64_t wid = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5 ;
    const int64_t nw = ((int64_t) gridDim.x * blockDim.x) >> 5 ;
    for (int64_t kk = wid ; kk < nvec ; kk += nw)
    {
        const int64_t lv = lpos [kk] ;
        if (lv < 0) continue ;
        const int64_t l0 = lp [lv], l1 = lp [lv+1] ;
        const int64_t shift = cum [kk] - pos [l0] ;
        for (int64_t e = l0 + lane ; e < l1 ; e += 32)
        {
            if (!flags [e]) continue ;
            const int64_t q = pos [e] + shift ;
            Ci [q] = li [e] ;
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
}

// =============================================================================================
// orchestration
// =============================================================================================
struct Bins
{
    DevBuf lists, counts ;
    int64_t n [NCLASS] ;
    const int32_t *list (int cl, int64_t nvec) const { return lists.as<int32_t> () + (int64_t) cl * nvec ; }
} ;

static gb200_status make_bins (const int64_t *flops, int64_t nvec, Bins &bins)
{
    Ctx &c = ctx () ;
    GB200_TRY (bins.lists.alloc ((size_t) NCLASS * (nvec > 0 ? nvec : 1) * sizeof (int32_t))) ;
    GB200_TRY (bins.counts.alloc (NCLASS * sizeof (unsigned int))) ;
    GB200_CUDA (cudaMemsetAsync (bins.counts.ptr, 0, NCLASS * sizeof (unsigned int), c.stream)) ;
    ClassLimits L ;
    for (int q = 0 ; q < 4 ; q++) L.lim [q] = class_limit [q] ;
    if (nvec > 0)
    {
        classify_kernel <<<grid_cap ((nvec + 255) / 256, 8), 256, 0, c.stream>>> (flops, nvec, L,
            bins.lists.as<int32_t> (), bins.counts.as<unsigned int> ()) ;
        count_launch () ;
    }
    GB200_CUDA (cudaMemcpyAsync (c.pinned, bins.counts.ptr, NCLASS * sizeof (unsigned int),
        cudaMemcpyDeviceToHost, c.stream)) ;
    GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
    for (int q = 0 ; q < NCLASS ; q++) bins.n [q] = ((unsigned int *) c.pinned) [q] ;
    return GB200_SUCCESS ;
}

struct HeavyWs
{
    DevBuf bitmap, rank, items, nitems ;
    int64_t nwords = 0, W = 0 ;
    bool persist = false ;          // the bitmaps of ALL heavy vectors fit: marked once, kept
} ;

static const int64_t HEAVY_TARGET = 1 << 16 ;               // flops per heavy work item, at most
static const int64_t HEAVY_TARGET_MIN = 1 << 12 ;           // ... and at least (small batches)

static gb200_status heavy_ws_alloc (HeavyWs &ws, int64_t vlen, int64_t nheavy, int64_t bnz,
    int64_t total_flops)
{
    ws.nwords = (vlen + 31) / 32 ;
    // bytes for bitmaps + ranks in flight: an eighth of the free HBM (180 GB per B200), at least 1 GiB
    size_t free_b = 0, total_b = 0 ;
    if (cudaMemGetInfo (&free_b, &total_b) != cudaSuccess) { cudaGetLastError () ; free_b = 0 ; }
    int64_t budget = (int64_t) (free_b / 8) ;
    if (budget < (1LL << 30)) budget = 1LL << 30 ;
    if (budget > (16LL << 30)) budget = 16LL << 30 ;
    int64_t W = budget / (8 * (ws.nwords > 0 ? ws.nwords : 1)) ;
    if (W < 1) W = 1 ;
    if (W >= nheavy) { W = nheavy ; ws.persist = true ; }
    ws.W = W ;
    GB200_TRY (ws.bitmap.alloc ((size_t) W * ws.nwords * 4)) ;
    GB200_TRY (ws.rank.alloc ((size_t) W * ws.nwords * 4)) ;
    // items per batch: sum of min (bjnz, ceil (flops/target)) <= min (bnz, total/target) + W
    int64_t cap = total_flops / HEAVY_TARGET_MIN + 1 ;
    if (cap > bnz) cap = bnz ;
    GB200_TRY (ws.items.alloc ((size_t) (cap + W + 1) * sizeof (HeavyItem))) ;
    GB200_TRY (ws.nitems.alloc (8)) ;
    return GB200_SUCCESS ;
}

static gb200_status heavy_make_items (HeavyWs &ws, const DMat &B, const int32_t *heavy, int64_t nh,
    const int64_t *flops, int64_t *nitems, int64_t wbase = 0, int64_t target = HEAVY_TARGET)
{
    Ctx &c = ctx () ;
    GB200_CUDA (cudaMemsetAsync (ws.nitems.ptr, 0, 8, c.stream)) ;
    heavy_items_kernel <<<grid_cap ((nh + 127) / 128, 8), 128, 0, c.stream>>> (B, heavy, nh, flops,
        target, wbase, ws.items.as<HeavyItem> (), ws.nitems.as<unsigned int> ()) ;
    count_launch () ;
    GB200_TRY (read_i64 (ws.nitems.as<int64_t> (), nitems)) ;
    *nitems &= 0xffffffffLL ;
    return GB200_SUCCESS ;
}

gb200_status run_saxpy (gb200_result_s *R, const gb200_dmatrix_s *Min, int mask_comp,
    const gb200_dmatrix_s *Ad, const gb200_dmatrix_s *Bd, const gb200_semiring &s)
{
    Ctx &c = ctx () ;
    const DMat &A = Ad->v ;
    const DMat &B = Bd->v ;
    const int64_t nvec = B.nvec ;
    const int64_t cvlen = A.vlen, cvdim = B.vdim ;
    // one GPU saxpy serves GUSTAVSON and HEAP requests alike; an explicit HEAP request is reported as
    // HEAP, as GB_AxB_select.c:139-143 would (its automatic heap choice, :95-128, is a CPU workspace
    // trade-off that has no analogue here and is reported as GUSTAVSON: INTEGRATION.md)
    R->info.method_used = (ctx ().method_request == GB200_METHOD_HEAP) ? GB200_METHOD_HEAP : GB200_METHOD_GUSTAVSON ;
    R->info.type_code = s.z_code ;

    // ---- mask policy of GB_AxB_sequential.c:76-95 ---------------------------------------------
    const gb200_dmatrix_s *M = Min ;
    if (M != nullptr && mask_comp) M = nullptr ;            // saxpy cannot use a complemented mask
    DevBuf flops, cum ;
    int64_t total = 0 ;
    if (M != nullptr && c.mask_policy == 2) M = nullptr ;   // the caller decided for all slices
    if (M != nullptr)
    {
        GB200_TRY (flopcount (&M->v, A, B, flops, cum, &total)) ;
        if (total <= M->v.nnz && c.mask_policy != 1) M = nullptr ;  // mask too dense to be worth using
    }
    if (M == nullptr) GB200_TRY (flopcount (nullptr, A, B, flops, cum, &total)) ;
    R->info.mask_applied = (M != nullptr) ? 1 : 0 ;
    R->info.flops = total ;
    const bool C_is_hyper = (cvdim > 1) &&
        (Ad->is_hyper_flag || Bd->is_hyper_flag || (M != nullptr && M->is_hyper_flag)) ;

    int acc_size = 0 ;
    const uint64_t ident = identity_bits (s.z_code, s.add_opcode, &acc_size) ;
    const int zsize = type_size (s.z_code) ;

    Bins bins ;
    GB200_TRY (make_bins (flops.as<int64_t> (), nvec, bins)) ;
    const int64_t nheavy = bins.n [CL_HEAVY] ;
    HeavyWs hws ;
    if (nheavy > 0) GB200_TRY (heavy_ws_alloc (hws, cvlen, nheavy, B.nnz, total)) ;

    SaxpyArgs sa ;
    memset (&sa, 0, sizeof (sa)) ;
    sa.A = A ; sa.B = B ;
    sa.mult_op = s.mult_opcode ; sa.flip = s.flipxy ;
    sa.bitmap = hws.bitmap.as<uint32_t> () ; sa.rank = hws.rank.as<int32_t> () ;
    sa.nwords = hws.nwords ;

    DevBuf Ci, Cx, ccum ;
    int64_t cnz = 0 ;

    if (M == nullptr)
    {
        // =====================================================================================
        // C = A*B : symbolic count -> scan -> symbolic fill -> numeric
        // =====================================================================================
        DevBuf cnt ;
        GB200_TRY (cnt.alloc ((nvec > 0 ? nvec : 1) * sizeof (int64_t))) ;
        GB200_CUDA (cudaMemsetAsync (cnt.ptr, 0, cnt.bytes, c.stream)) ;
        // small index range: a shared-memory bitmap replaces the hash set (and the sort of the fill
        // pass) wherever scanning vlen/32 words costs less than sorting the vector
        const int symb_words = (int) ((cvlen + 31) / 32) ;
        const char *benv = getenv ("GB200_SYM_BITMAP") ;
        const bool symb_ok = (cvlen <= SYMB_MAX_VLEN) && !(benv != nullptr && atoi (benv) == 0) ;
        auto use_bitmap = [&] (int cl) { return symb_ok && symb_words <= 8 * class_limit [cl] ; } ;
        if (symb_ok && (size_t) symb_words * 4 > 48 * 1024)
        {
            GB200_CUDA (cudaFuncSetAttribute (sym_bitmap_kernel<false>,
                cudaFuncAttributeMaxDynamicSharedMemorySize, symb_words * 4)) ;
            GB200_CUDA (cudaFuncSetAttribute (sym_bitmap_kernel<true>,
                cudaFuncAttributeMaxDynamicSharedMemorySize, symb_words * 4)) ;
        }
        for (int cl = 0 ; cl < 4 ; cl++)
        {
            if (bins.n [cl] == 0) continue ;
            if (use_bitmap (cl))
            {
                sym_bitmap_kernel<false> <<<grid_cap (bins.n [cl], 32), class_threads [cl],
                    (size_t) symb_words * 4, c.stream>>> (A, B, bins.list (cl, nvec), bins.n [cl],
                    symb_words, cnt.as<int64_t> (), nullptr, nullptr) ;
                count_launch () ;
                continue ;
            }
            const size_t smem = (size_t) 4 << class_log [cl] ;
            if (smem > 48 * 1024)
                GB200_CUDA (cudaFuncSetAttribute (sym_hash_kernel<false>,
                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) ;
            sym_hash_kernel<false> <<<grid_cap (bins.n [cl], 32), class_threads [cl], smem, c.stream>>> (
                A, B, bins.list (cl, nvec), bins.n [cl], class_log [cl], cnt.as<int64_t> (), nullptr, nullptr) ;
            count_launch () ;
        }
        for (int64_t h0 = 0 ; h0 < nheavy ; h0 += hws.W)
        {
            const int64_t nh = (nheavy - h0 < hws.W) ? (nheavy - h0) : hws.W ;
            const int32_t *heavy = bins.list (CL_HEAVY, nvec) + h0 ;
            int64_t nitems = 0 ;
            GB200_TRY (heavy_make_items (hws, B, heavy, nh, flops.as<int64_t> (), &nitems)) ;
            GB200_CUDA (cudaMemsetAsync (hws.bitmap.ptr, 0, (size_t) nh * hws.nwords * 4, c.stream)) ;
            heavy_mark_kernel <<<grid_cap (nitems, 16), 256, 0, c.stream>>> (A, B,
                hws.items.as<HeavyItem> (), nitems, hws.bitmap.as<uint32_t> (), hws.nwords) ;
            heavy_count_kernel <<<grid_cap (nh, 8), 256, 0, c.stream>>> (heavy, nh,
                hws.bitmap.as<uint32_t> (), hws.nwords, cnt.as<int64_t> ()) ;
            count_launch (2) ;
        }
        GB200_CUDA (cudaGetLastError ()) ;
        GB200_TRY (ccum.alloc ((nvec + 1) * sizeof (int64_t))) ;
        GB200_TRY (scan_i64 (cnt.as<int64_t> (), ccum.as<int64_t> (), nvec)) ;
        GB200_TRY (read_i64 (ccum.as<int64_t> () + nvec, &cnz)) ;

        GB200_TRY (Ci.alloc ((cnz > 0 ? cnz : 1) * sizeof (int32_t))) ;
        DevBuf acc ;
        GB200_TRY (acc.alloc ((size_t) (cnz > 0 ? cnz : 1) * acc_size)) ;
        GB200_TRY (fill_bits (acc.ptr, acc_size, ident, cnz)) ;

        sa.lp = ccum.as<int64_t> () ; sa.li = Ci.as<int32_t> () ; sa.lpos = nullptr ;
        sa.acc = acc.ptr ; sa.flags = nullptr ; sa.masked = 0 ;

        // Vectors whose pattern fits a shared-memory table (<= 4096 entries): fused fill + numeric with
        // a shared-memory hash accumulator, binned by entry count.  GB200_SAXPY_HASH=0: the two-kernel
        // path (pattern first, then a binary search and a global atomic per product) for all of them.
        // Measured (profiles/r2, one B200): Erdos-Renyi 2^20 10.6 -> 8.0 ms with the two small classes
        // fused; the third class (<= 4096 entries: a 128 KB table, one block per SM) loses to the
        // two-kernel path on RMAT 18 (84 -> 102 ms), so it is off unless GB200_SAXPY_HASH=2.
        const char *henv = getenv ("GB200_SAXPY_HASH") ;
        const bool fused = !(henv != nullptr && atoi (henv) == 0) ;
        const bool fuse_mid = (henv != nullptr && atoi (henv) == 2) ;
        const int64_t hash_limit [3] = { 128, 1024, fuse_mid ? 4096 : 1024 } ;
        static const int hash_log [3] = { 8, 11, 13 } ;
        static const int hash_threads [3] = { 32, 128, 512 } ;
        Bins cb ;                                   // lists 0..2: fused classes; 3: pattern too long
        if (fused)
        {
            GB200_TRY (cb.lists.alloc ((size_t) NCLASS * (nvec > 0 ? nvec : 1) * sizeof (int32_t))) ;
            GB200_TRY (cb.counts.alloc (NCLASS * sizeof (unsigned int))) ;
            GB200_CUDA (cudaMemsetAsync (cb.counts.ptr, 0, NCLASS * sizeof (unsigned int), c.stream)) ;
            CntLimits CL ;
            for (int q = 0 ; q < 3 ; q++) CL.lim [q] = hash_limit [q] ;
            CL.mid_flops = class_limit [2] ; CL.heavy_flops = class_limit [3] ;
            if (nvec > 0)
            {
                classify_cnt_kernel <<<grid_cap ((nvec + 255) / 256, 8), 256, 0, c.stream>>> (
                    flops.as<int64_t> (), ccum.as<int64_t> (), nvec, CL, cb.lists.as<int32_t> (),
                    cb.counts.as<unsigned int> ()) ;
                count_launch () ;
            }
            GB200_CUDA (cudaMemcpyAsync (c.pinned, cb.counts.ptr, NCLASS * sizeof (unsigned int),
                cudaMemcpyDeviceToHost, c.stream)) ;
            GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
            for (int q = 0 ; q < NCLASS ; q++) cb.n [q] = ((unsigned int *) c.pinned) [q] ;
            sa.Ci_out = Ci.as<int32_t> () ;
            const int64_t heavy_nwords = sa.nwords ;
            for (int cl = 0 ; cl < 3 ; cl++)
            {
                if (cb.n [cl] == 0) continue ;
                sa.cols = cb.list (cl, nvec) ; sa.ncols = cb.n [cl] ; sa.hash_log = hash_log [cl] ;
                // a vlen-bit bitmap next to the table replaces the sort wherever scanning it costs less
                // than sorting the vector (same rule as the pattern-only kernels above)
                sa.nwords = use_bitmap (cl == 0 ? 0 : cl) ? symb_words : 0 ;
                // the smallest class without a bitmap: a warp per vector, three warps per block
                const char *hw_env = getenv ("GB200_SAXPY_HASH_WARP") ;
                const bool warp_class = (cl == 0 && sa.nwords == 0 && hash_log [0] == HASHW_LOG
                    && !(hw_env != nullptr && atoi (hw_env) == 0)) ;
                const bool ok = warp_class
                    ? launch_typed (s.xy_code, FAM_SAXPY_HASH_WARP, s.z_code, s.add_opcode, s.mult_opcode, &sa,
                        grid_cap ((cb.n [cl] + HASHW_WARPS - 1) / HASHW_WARPS, 16), 32 * HASHW_WARPS)
                    : launch_typed (s.xy_code, FAM_SAXPY_HASH, s.z_code, s.add_opcode, s.mult_opcode, &sa,
                        grid_cap (cb.n [cl], (cl == 0) ? 32 : ((cl == 1) ? 8 : 1)), hash_threads [cl]) ;
                if (!ok) { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
            }
            sa.nwords = heavy_nwords ;
            for (int q = 3 ; q <= 4 ; q++)
            {
                // longer patterns of the two upper shared-memory flop classes: pattern first (bitmap, or
                // hash set + sort), then the slot search per product
                if (cb.n [q] == 0) continue ;
                const int cl = q - 1 ;              // flop class 2 or 3
                if (use_bitmap (cl))
                {
                    sym_bitmap_kernel<true> <<<grid_cap (cb.n [q], 32), class_threads [cl],
                        (size_t) symb_words * 4, c.stream>>> (A, B, cb.list (q, nvec), cb.n [q],
                        symb_words, nullptr, ccum.as<int64_t> (), Ci.as<int32_t> ()) ;
                }
                else
                {
                    const size_t smem = ((size_t) 4 << class_log [cl]) + ((size_t) 2 << class_log [cl]) ;
                    if (smem > 48 * 1024)
                        GB200_CUDA (cudaFuncSetAttribute (sym_hash_kernel<true>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) ;
                    sym_hash_kernel<true> <<<grid_cap (cb.n [q], 32), class_threads [cl], smem, c.stream>>> (
                        A, B, cb.list (q, nvec), cb.n [q], class_log [cl], nullptr, ccum.as<int64_t> (),
                        Ci.as<int32_t> ()) ;
                }
                count_launch () ;
                sa.cols = cb.list (q, nvec) ; sa.ncols = cb.n [q] ;
                if (!launch_typed (s.xy_code, FAM_SAXPY_LIGHT, s.z_code, s.add_opcode, s.mult_opcode, &sa,
                    grid_cap (cb.n [q], 32), class_threads [cl]))
                { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
            }
        }
        else
        for (int cl = 0 ; cl < 4 ; cl++)
        {
            if (bins.n [cl] == 0) continue ;
            if (use_bitmap (cl))
            {
                sym_bitmap_kernel<true> <<<grid_cap (bins.n [cl], 32), class_threads [cl],
                    (size_t) symb_words * 4, c.stream>>> (A, B, bins.list (cl, nvec), bins.n [cl],
                    symb_words, nullptr, ccum.as<int64_t> (), Ci.as<int32_t> ()) ;
            }
            else
            {
                const size_t smem = ((size_t) 4 << class_log [cl]) + ((size_t) 2 << class_log [cl]) ;
                if (smem > 48 * 1024)
                    GB200_CUDA (cudaFuncSetAttribute (sym_hash_kernel<true>,
                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)) ;
                sym_hash_kernel<true> <<<grid_cap (bins.n [cl], 32), class_threads [cl], smem, c.stream>>> (
                    A, B, bins.list (cl, nvec), bins.n [cl], class_log [cl], nullptr, ccum.as<int64_t> (),
                    Ci.as<int32_t> ()) ;
            }
            count_launch () ;
            sa.cols = bins.list (cl, nvec) ; sa.ncols = bins.n [cl] ;
            if (!launch_typed (s.xy_code, FAM_SAXPY_LIGHT, s.z_code, s.add_opcode, s.mult_opcode, &sa,
                grid_cap (bins.n [cl], 32), class_threads [cl]))
            { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
        }
        // The accumulators of the heavy vectors in flight are hit by one atomic per product: a batch
        // is cut so that they (and the bitmaps and ranks) stay resident in L2, otherwise every atomic
        // is a DRAM read-modify-write
        std::vector<int64_t> hcnt ((size_t) nheavy), hflops ((size_t) nheavy) ;
        if (nheavy > 0)
        {
            DevBuf hc ;
            GB200_TRY (hc.alloc (2 * nheavy * sizeof (int64_t))) ;
            gather_i64_kernel <<<grid_cap ((nheavy + 255) / 256, 8), 256, 0, c.stream>>> (cnt.as<int64_t> (),
                bins.list (CL_HEAVY, nvec), nheavy, hc.as<int64_t> ()) ;
            gather_i64_kernel <<<grid_cap ((nheavy + 255) / 256, 8), 256, 0, c.stream>>> (flops.as<int64_t> (),
                bins.list (CL_HEAVY, nvec), nheavy, hc.as<int64_t> () + nheavy) ;
            count_launch (2) ;
            GB200_CUDA (cudaMemcpyAsync (hcnt.data (), hc.ptr, nheavy * sizeof (int64_t),
                cudaMemcpyDeviceToHost, c.stream)) ;
            GB200_CUDA (cudaMemcpyAsync (hflops.data (), hc.as<int64_t> () + nheavy,
                nheavy * sizeof (int64_t), cudaMemcpyDeviceToHost, c.stream)) ;
            GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
        }
        const char *l2env = getenv ("GB200_HEAVY_L2_MB") ;
        const int64_t l2_budget = ((l2env != nullptr && atoll (l2env) > 0) ? atoll (l2env) : 128) << 20 ;
        if (hws.persist && nheavy > 0)
        {
            // Every heavy bitmap is still there from the symbolic phase: rank them all in one launch,
            // lay the work items of all vectors out in list order, and then only the numeric kernel
            // runs per L2-sized batch -- no host round trip between batches.
            const int32_t *heavy = bins.list (CL_HEAVY, nvec) ;
            heavy_rank_kernel <<<grid_cap (nheavy, 8), 256, 0, c.stream>>> (heavy, nheavy,
                hws.bitmap.as<uint32_t> (), hws.rank.as<int32_t> (), hws.nwords, ccum.as<int64_t> (),
                Ci.as<int32_t> ()) ;
            count_launch () ;
            std::vector<int64_t> bstart ;
            int64_t bytes = 0, hflops_total = 0 ;
            for (int64_t t = 0 ; t < nheavy ; t++)
            {
                const int64_t add = hcnt [t] * acc_size + hws.nwords * 8 ;
                if (t == 0 || bytes + add > l2_budget) { bstart.push_back (t) ; bytes = 0 ; }
                bytes += add ;
                hflops_total += hflops [t] ;
            }
            bstart.push_back (nheavy) ;
            const int64_t nbatch = (int64_t) bstart.size () - 1 ;
            // enough work items per batch to fill the machine
            int64_t target = hflops_total / (nbatch * (int64_t) c.sm_count * 16) ;
            if (target > HEAVY_TARGET) target = HEAVY_TARGET ;
            if (target < HEAVY_TARGET_MIN) target = HEAVY_TARGET_MIN ;
            DevBuf nch, ioff ;
            GB200_TRY (nch.alloc (nheavy * sizeof (int64_t))) ;
            GB200_TRY (ioff.alloc ((nheavy + 1) * sizeof (int64_t))) ;
            heavy_items_ordered_kernel <<<grid_cap ((nheavy + 127) / 128, 8), 128, 0, c.stream>>> (B, heavy,
                nheavy, flops.as<int64_t> (), target, nch.as<int64_t> (), nullptr, nullptr) ;
            count_launch () ;
            GB200_TRY (scan_i64 (nch.as<int64_t> (), ioff.as<int64_t> (), nheavy)) ;
            heavy_items_ordered_kernel <<<grid_cap ((nheavy + 127) / 128, 8), 128, 0, c.stream>>> (B, heavy,
                nheavy, flops.as<int64_t> (), target, nullptr, ioff.as<int64_t> (),
                hws.items.as<HeavyItem> ()) ;
            count_launch () ;
            std::vector<int64_t> hoff ((size_t) nheavy + 1) ;
            GB200_CUDA (cudaMemcpyAsync (hoff.data (), ioff.ptr, (nheavy + 1) * sizeof (int64_t),
                cudaMemcpyDeviceToHost, c.stream)) ;
            GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
            for (int64_t b = 0 ; b < nbatch ; b++)
            {
                const int64_t i0 = hoff [bstart [b]], i1 = hoff [bstart [b+1]] ;
                if (i1 <= i0) continue ;
                sa.items = hws.items.as<HeavyItem> () + i0 ; sa.nitems = i1 - i0 ;
                if (!launch_typed (s.xy_code, FAM_SAXPY_HEAVY, s.z_code, s.add_opcode, s.mult_opcode, &sa,
                    grid_cap (i1 - i0, 16), 256))
                { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
            }
        }
        else for (int64_t h0 = 0, nh = 0 ; h0 < nheavy ; h0 += nh)
        {
            int64_t bytes = 0, bflops = 0 ;
            for (nh = 0 ; h0 + nh < nheavy && nh < hws.W ; nh++)
            {
                const int64_t add = hcnt [h0 + nh] * acc_size + hws.nwords * 8 ;
                if (nh > 0 && bytes + add > l2_budget) break ;
                bytes += add ;
                bflops += hflops [h0 + nh] ;
            }
            // enough work items to fill the machine even when the batch is small
            int64_t target = bflops / ((int64_t) c.sm_count * 16) ;
            if (target > HEAVY_TARGET) target = HEAVY_TARGET ;
            if (target < HEAVY_TARGET_MIN) target = HEAVY_TARGET_MIN ;
            const int32_t *heavy = bins.list (CL_HEAVY, nvec) + h0 ;
            const int64_t wbase = hws.persist ? h0 : 0 ;        // bitmap of the batch's first vector
            int64_t nitems = 0 ;
            GB200_TRY (heavy_make_items (hws, B, heavy, nh, flops.as<int64_t> (), &nitems, wbase, target)) ;
            if (!hws.persist)
            {
                GB200_CUDA (cudaMemsetAsync (hws.bitmap.ptr, 0, (size_t) nh * hws.nwords * 4, c.stream)) ;
                heavy_mark_kernel <<<grid_cap (nitems, 16), 256, 0, c.stream>>> (A, B,
                    hws.items.as<HeavyItem> (), nitems, hws.bitmap.as<uint32_t> (), hws.nwords) ;
                count_launch () ;
            }
            heavy_rank_kernel <<<grid_cap (nh, 8), 256, 0, c.stream>>> (heavy, nh,
                hws.bitmap.as<uint32_t> () + wbase * hws.nwords, hws.rank.as<int32_t> () + wbase * hws.nwords,
                hws.nwords, ccum.as<int64_t> (), Ci.as<int32_t> ()) ;
            count_launch () ;
            sa.items = hws.items.as<HeavyItem> () ; sa.nitems = nitems ;
            if (!launch_typed (s.xy_code, FAM_SAXPY_HEAVY, s.z_code, s.add_opcode, s.mult_opcode, &sa,
                grid_cap (nitems, 16), 256))
            { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
        }
        GB200_CUDA (cudaGetLastError ()) ;
        if (zsize >= 4) Cx = std::move (acc) ;
        else
        {
            GB200_TRY (Cx.alloc ((size_t) (cnz > 0 ? cnz : 1) * zsize)) ;
            GB200_TRY (convert_acc (acc.ptr, acc_size, Cx.ptr, s.z_code, cnz)) ;
        }
    }
    else
    {
        // =====================================================================================
        // C<M> = A*B : the slot list is the (structural) mask itself
        // =====================================================================================
        DMat Mv ; DevBuf Mp2, Mi2 ;
        GB200_TRY (filter_mask (M, Mv, Mp2, Mi2)) ;
        const int64_t mnz = Mv.nnz ;
        DevBuf lpos, acc, flags ;
        GB200_TRY (lpos.alloc ((nvec > 0 ? nvec : 1) * sizeof (int64_t))) ;
        GB200_TRY (launch_mask_pos (B, Mv, lpos.as<int64_t> ())) ;
        GB200_TRY (acc.alloc ((size_t) (mnz > 0 ? mnz : 1) * acc_size)) ;
        GB200_TRY (fill_bits (acc.ptr, acc_size, ident, mnz)) ;
        GB200_TRY (flags.alloc (mnz > 0 ? mnz : 1)) ;
        GB200_CUDA (cudaMemsetAsync (flags.ptr, 0, flags.bytes, c.stream)) ;

        sa.lp = Mv.p ; sa.li = Mv.i ; sa.lpos = lpos.as<int64_t> () ;
        sa.acc = acc.ptr ; sa.flags = flags.as<uint8_t> () ; sa.masked = 1 ;
        for (int cl = 0 ; cl < 4 ; cl++)
        {
            if (bins.n [cl] == 0) continue ;
            sa.cols = bins.list (cl, nvec) ; sa.ncols = bins.n [cl] ;
            if (!launch_typed (s.xy_code, FAM_SAXPY_LIGHT, s.z_code, s.add_opcode, s.mult_opcode, &sa,
                grid_cap (bins.n [cl], 32), class_threads [cl]))
            { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
        }
        for (int64_t h0 = 0 ; h0 < nheavy ; h0 += hws.W)
        {
            const int64_t nh = (nheavy - h0 < hws.W) ? (nheavy - h0) : hws.W ;
            const int32_t *heavy = bins.list (CL_HEAVY, nvec) + h0 ;
            int64_t nitems = 0 ;
            GB200_TRY (heavy_make_items (hws, B, heavy, nh, flops.as<int64_t> (), &nitems)) ;
            GB200_CUDA (cudaMemsetAsync (hws.bitmap.ptr, 0, (size_t) nh * hws.nwords * 4, c.stream)) ;
            heavy_mark_list_kernel <<<grid_cap (nh, 8), 256, 0, c.stream>>> (heavy, nh, Mv.p, Mv.i,
                lpos.as<int64_t> (), hws.bitmap.as<uint32_t> (), hws.nwords) ;
            heavy_rank_kernel <<<grid_cap (nh, 8), 256, 0, c.stream>>> (heavy, nh,
                hws.bitmap.as<uint32_t> (), hws.rank.as<int32_t> (), hws.nwords, nullptr, nullptr) ;
            count_launch (2) ;
            sa.items = hws.items.as<HeavyItem> () ; sa.nitems = nitems ;
            if (!launch_typed (s.xy_code, FAM_SAXPY_HEAVY, s.z_code, s.add_opcode, s.mult_opcode, &sa,
                grid_cap (nitems, 16), 256))
            { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
        }
        GB200_CUDA (cudaGetLastError ()) ;
        // compaction of the slots that received at least one product, in mask order
        DevBuf pos, cnt ;
        GB200_TRY (pos.alloc ((mnz + 1) * sizeof (int64_t))) ;
        GB200_TRY (scan_u8 (flags.as<uint8_t> (), pos.as<int64_t> (), mnz)) ;
        GB200_TRY (cnt.alloc ((nvec > 0 ? nvec : 1) * sizeof (int64_t))) ;
        if (nvec > 0)
        {
            masked_counts_kernel <<<grid_cap ((nvec + 255) / 256, 8), 256, 0, c.stream>>> (
                lpos.as<int64_t> (), Mv.p, pos.as<int64_t> (), nvec, cnt.as<int64_t> ()) ;
            count_launch () ;
        }
        GB200_TRY (ccum.alloc ((nvec + 1) * sizeof (int64_t))) ;
        GB200_TRY (scan_i64 (cnt.as<int64_t> (), ccum.as<int64_t> (), nvec)) ;
        GB200_TRY (read_i64 (ccum.as<int64_t> () + nvec, &cnz)) ;
        GB200_TRY (Ci.alloc ((cnz > 0 ? cnz : 1) * sizeof (int32_t))) ;
        GB200_TRY (Cx.alloc ((size_t) (cnz > 0 ? cnz : 1) * zsize)) ;
        if (nvec > 0 && cnz > 0)
        {
            masked_gather_kernel <<<grid_cap ((nvec + 7) / 8, 8), 256, 0, c.stream>>> (
                lpos.as<int64_t> (), Mv.p, Mv.i, flags.as<uint8_t> (), pos.as<int64_t> (),
                ccum.as<int64_t> (), acc.ptr, acc_size, zsize, s.z_code == GB200_BOOL, nvec,
                Ci.as<int32_t> (), Cx.ptr) ;
            count_launch () ;
        }
        GB200_CUDA (cudaGetLastError ()) ;
        // everything above is stream-ordered; buffers local to this scope are freed in order
    }
    return assemble (R, nvec, B.hyper ? B.h : nullptr, B.hyper != 0, ccum, Ci, Cx, cnz, C_is_hyper,
        cvlen, cvdim) ;
}

} // namespace gb200
