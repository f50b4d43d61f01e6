// engine_dot.cu -- C<M>=A'*B, C<!M>=A'*B and C=A'*B by dot products: the GPU replacement of
// GB_AxB_dot (reference Source/GB_AxB_dot.c:39-317 and Source/Template/GB_AxB_dot_{mask,compmask,
// nomask,cij}.c), plus the entry points gb200_AxB_device / gb200_AxB_host.
//
// Every candidate C(i,j) is a *pair*; a group of lanes computes the pair (dot_kernel, kernels.cuh),
// leaving a value and a "some index matched" flag per pair; a scan of the flags then packs the
// results.  Pairs are enumerated in (j, i) order, so packed indices are ascending in every vector.
#include <chrono>
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

// stored-vector position of every entry of M: a warp takes 32 consecutive vectors; a lane fills the
// run of a short vector itself, the warp fills the long ones together
__global__ void expand_vec_kernel (const int64_t *__restrict__ p, int64_t nvec, int64_t nnz,
    int32_t *__restrict__ mvec)
{
    const int lane = threadIdx.x & 31 ;
    const int64_t wid = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
    const int64_t nw = ((int64_t) gridDim.x * blockDim.x) >> 5 ;
    for (int64_t v0 = wid * 32 ; v0 < nvec ; v0 += nw * 32)
    {
        const int64_t v = v0 + lane ;
        int64_t e0 = 0, e1 = 0 ;
        if (v < nvec) { e0 = __ldg (p + v) ; e1 = __ldg (p + v + 1) ; }
        const bool is_long = (e1 - e0 > 64) ;
        if (!is_long) for (int64_t e = e0 ; e < e1 ; e++) mvec [e] = (int32_t) v ;
        unsigned todo = __ballot_sync (0xffffffffu, is_long) ;
        while (todo)
        {
            const int src = __ffs (todo) - 1 ;
            todo &= todo - 1 ;
            const int64_t s0 = __shfl_sync (0xffffffffu, e0, src), s1 = __shfl_sync (0xffffffffu, e1, src) ;
            for (int64_t e = s0 + lane ; e < s1 ; e += 32) mvec [e] = (int32_t) (v0 + src) ;
        }
    }
    (void) nnz ;
}

// pack the flagged pairs.  rows: li[e] (mask mode) or the name of A's (e % anvec)-th vector.
__global__ void dot_gather_kernel (const uint8_t *__restrict__ flags, const int64_t *__restrict__ pos,
    int64_t npairs, const int32_t *__restrict__ li, DMat A, const void *__restrict__ acc, int acc_size,
    int zsize, int is_bool, int32_t *__restrict__ Ci, void *__restrict__ Cx)
{
    for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; e < npairs ;
        e += (int64_t) gridDim.x * blockDim.x)
    {
        if (!flags [e]) continue ;
        const int64_t q = pos [e] ;
        Ci [q] = li ? li [e] : (int32_t) dm_vecname (A, e % A.nvec) ;
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

// cum[v] = pos[p[v]] for list-shaped pair spaces; cnt[jb] for rectangular ones
__global__ void dot_cum_list_kernel (const int64_t *__restrict__ p, const int64_t *__restrict__ pos,
    int64_t nvec, int64_t *__restrict__ cum)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t <= nvec ;
        t += (int64_t) gridDim.x * blockDim.x) cum [t] = pos [p [t]] ;
}

__global__ void dot_cnt_rect_kernel (const int64_t *__restrict__ pos, int64_t anvec, int64_t nb,
    int64_t *__restrict__ cnt)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < nb ;
        t += (int64_t) gridDim.x * blockDim.x) cnt [t] = pos [(t + 1) * anvec] - pos [t * anvec] ;
}

// tasks per work item of dotg_kernel (regular owners): every item of an owner builds the owner's table
// again, so items are large (measured, tri scale 22: 128: 46.5, 256: 43.8, 512: 42.8, 1024: 42.4 ms)
constexpr int64_t DOTG_CHUNK = 1024 ;

// Per stored vector: where it starts, how long it is, its first and last index -- one record, so that
// the classification of a pair costs one random sector per side instead of three (pointer pair, first
// index, last index).  16 bytes while positions fit 32 bits (two records per sector, and the table of a
// 4 M-vector operand is 67 MB: it stays in L2), 32 bytes otherwise.
struct VecRec { int64_t p0 ; int64_t len ; int32_t first ; int32_t last ; } ;
struct __align__ (16) VecInfo16
{
    uint32_t p0 ; uint32_t len ; int32_t first ; int32_t last ;
    __device__ __forceinline__ static void store (VecInfo16 *q, const VecRec &v)
    {
        *(int4 *) q = make_int4 ((int) (uint32_t) v.p0, (int) (uint32_t) v.len, v.first, v.last) ;
    }
    __device__ __forceinline__ static VecRec load (const VecInfo16 *__restrict__ q)
    {
        const int4 a = __ldg ((const int4 *) q) ;
        VecRec v ;
        v.p0 = (int64_t) (uint32_t) a.x ; v.len = (int64_t) (uint32_t) a.y ; v.first = a.z ; v.last = a.w ;
        return v ;
    }
} ;
struct __align__ (32) VecInfo32
{
    int64_t p0 ; int64_t len ; int32_t first ; int32_t last ; int64_t pad ;
    __device__ __forceinline__ static void store (VecInfo32 *q, const VecRec &v)
    {
        VecInfo32 w ;
        w.p0 = v.p0 ; w.len = v.len ; w.first = v.first ; w.last = v.last ; w.pad = 0 ;
        *q = w ;
    }
    __device__ __forceinline__ static VecRec load (const VecInfo32 *__restrict__ q)
    {
        const int4 a = __ldg ((const int4 *) q), b = __ldg (((const int4 *) q) + 1) ;
        VecRec v ;
        v.p0 = ((int64_t) (uint32_t) a.x) | ((int64_t) a.y << 32) ;
        v.len = ((int64_t) (uint32_t) a.z) | ((int64_t) a.w << 32) ;
        v.first = b.x ; v.last = b.y ;
        return v ;
    }
} ;

template <class VI>
__global__ void vec_info_kernel (DMat X, VI *__restrict__ info)
{
    for (int64_t k = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; k < X.nvec ;
        k += (int64_t) gridDim.x * blockDim.x)
    {
        VecRec v ;
        v.p0 = X.p [k] ; v.len = X.p [k+1] - v.p0 ;
        v.first = (v.len > 0) ? __ldg (X.i + v.p0) : 0 ;
        v.last = (v.len > 0) ? __ldg (X.i + v.p0 + v.len - 1) : 0 ;
        VI::store (info + k, v) ;
    }
}

// Both lists are sorted, so a match can only lie between the owner's first and last index: the walked
// list is trimmed to that range before it becomes a task (at most two binary searches per pair).  For
// C<L>=L*U' every pair keeps only the indices between j and i: 43 % fewer probes on RMAT graphs.
// wfirst / wlast: the walked list's own first and last index (a side whose end is already inside the
// range needs no search: the common case for one side).
__device__ __forceinline__ void dotg_trim (const int32_t *__restrict__ Wi, int64_t w0, int64_t w1,
    int32_t wfirst, int32_t wlast, int32_t omin, int32_t omax, int64_t &t0, int64_t &t1)
{
    int64_t l = w0, h = w1 ;
    if (wfirst < omin)
    {
        while (l < h)
        {
            const int64_t mid = (l + h) >> 1 ;
            if (__ldg (Wi + mid) < omin) l = mid + 1 ; else h = mid ;
        }
    }
    t0 = l ;
    h = w1 ;
    if (l < w1 && wlast > omax)
    {
        while (l < h)
        {
            const int64_t mid = (l + h) >> 1 ;
            if (__ldg (Wi + mid) <= omax) l = mid + 1 ; else h = mid ;
        }
        t1 = l ;
    }
    else t1 = w1 ;
}

// ---- set-up of the masked dot: two passes over the mask entries -------------------------------------
// A pair is DEAD (an empty side, or nothing left after the trim), SMALL (owner shorter than DOTG_SMALL:
// dot_kernel, no table), B-OWNED (B(:,j) is the longer list: A(:,i) is walked) or A-OWNED.  B-owned
// pairs keep the mask's order, so the tasks of an owner B(:,j) are a contiguous run found by a scan;
// A-owned pairs are regrouped by i with a counting sort (counts by atomics in the first pass, cursors in
// the second).  A TASK is one pair, or one DOTG_SEG-long piece of a pair with a longer walk.
enum { PK_DEAD = 0, PK_SMALL = 1, PK_BOWN = 2, PK_AOWN = 3 } ;

// pass 1: classification with the trim.  w0 [e] = where the (trimmed) walk starts in the walked matrix,
// lk [e] = its length | kind << 30, nt0 [e] = tasks of a B-owned pair (else 0), cntA [ka] += tasks of an
// A-owned pair, slist = the small pairs (any order: they are independent).
// trim: 0 walk whole lists; 1 (default) always search; 2 no search in a walked list of at most 32 indices:
// it is one row of the walk with or without the trim -- such a pair is only dropped when the two index
// ranges do not meet (measured: the same step time, but 28 % more DRAM traffic in the walk, because an
// untrimmed row touches more sectors).
template <class VI>
__global__ void dotg_classify_kernel (DMat A, DMat B, DMat M, const VI *__restrict__ infoA,
    const VI *__restrict__ infoB, const int32_t *__restrict__ mvec,
    int64_t mnz, int trim, int64_t *__restrict__ w0out, int32_t *__restrict__ lk, int32_t *__restrict__ nt0,
    unsigned long long *__restrict__ cntA, int32_t *__restrict__ slist, unsigned int *__restrict__ nsmall)
{
    const int lane = threadIdx.x & 31 ;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x ;
    const int64_t niter = (mnz + stride - 1) / stride ;
    int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ;
    for (int64_t itn = 0 ; itn < niter ; itn++, e += stride)
    {
        int kind = PK_DEAD ;
        int32_t len = 0, ntk = 0 ;
        int64_t wstart = 0 ;
        if (e < mnz)
        {
            const int64_t ka = dm_vecpos (A, M.i [e]) ;
            const int64_t kb = dm_vecpos (B, dm_vecname (M, mvec [e])) ;
            if (ka >= 0 && kb >= 0)
            {
                const VecRec va = VI::load (infoA + ka), vb = VI::load (infoB + kb) ;
                const int64_t ainz = va.len, bjnz = vb.len ;
                if (ainz > 0 && bjnz > 0)
                {
                    const bool walkA = dot_walkA (ainz, bjnz, A.vlen) ;
                    const int64_t olen = walkA ? bjnz : ainz ;
                    if (olen < DOTG_SMALL) kind = PK_SMALL ;
                    else
                    {
                        const VecRec &vw = walkA ? va : vb ;        // walked
                        const VecRec &vo = walkA ? vb : va ;        // owner
                        int64_t t0 = vw.p0, t1 = vw.p0 + vw.len ;
                        if (trim && (vw.last < vo.first || vw.first > vo.last)) t1 = t0 ;
                        else if (trim == 1 || (trim == 2 && vw.len > 32))
                            dotg_trim (walkA ? A.i : B.i, vw.p0, vw.p0 + vw.len, vw.first, vw.last,
                                vo.first, vo.last, t0, t1) ;
                        len = (int32_t) (t1 - t0) ;
                        wstart = t0 ;
                        if (len > 0)
                        {
                            ntk = (len + DOTG_SEG - 1) / DOTG_SEG ;
                            if (walkA) kind = PK_BOWN ;
                            else { kind = PK_AOWN ; atomicAdd (cntA + ka, (unsigned long long) ntk) ; }
                        }
                    }
                }
            }
            w0out [e] = wstart ;
            lk [e] = len | (kind << 30) ;
            nt0 [e] = (kind == PK_BOWN) ? ntk : 0 ;
        }
        // the small pairs of the warp are appended with one atomic
        const unsigned sm = __ballot_sync (0xffffffffu, kind == PK_SMALL) ;
        if (sm)
        {
            unsigned int base = 0 ;
            if (lane == 0) base = atomicAdd (nsmall, (unsigned int) __popc (sm)) ;
            base = __shfl_sync (0xffffffffu, base, 0) ;
            if (kind == PK_SMALL) slist [base + __popc (sm & ((1u << lane) - 1u))] = (int32_t) e ;
        }
    }
}

// pass 2: the task records.  B-owned pairs: at toff0 [e]; A-owned pairs: behind them, grouped by the
// vector of A (offA = scan of cntA, curA = running cursor of every vector).
__global__ void dotg_scatter_kernel (DMat A, DMat M, int64_t mnz, const int64_t *__restrict__ w0in,
    const int32_t *__restrict__ lk, const int64_t *__restrict__ toff0, const int64_t *__restrict__ offA,
    unsigned long long *__restrict__ curA, int64_t ntask0, DotTask *__restrict__ tasks)
{
    for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; e < mnz ;
        e += (int64_t) gridDim.x * blockDim.x)
    {
        const int32_t v = lk [e] ;
        const int kind = (int) (((uint32_t) v) >> 30) ;
        if (kind < PK_BOWN) continue ;
        const int32_t w = v & 0x3fffffff ;
        const int32_t ntk = (w + DOTG_SEG - 1) / DOTG_SEG ;
        int64_t q ;
        if (kind == PK_BOWN) q = toff0 [e] ;
        else
        {
            const int64_t ka = dm_vecpos (A, M.i [e]) ;
            q = ntask0 + offA [ka] + (int64_t) atomicAdd (curA + ka, (unsigned long long) ntk) ;
        }
        const int64_t ws = w0in [e] ;
        const bool split = (w > DOTG_SEG) ;
        for (int32_t s0 = 0 ; s0 < w ; s0 += DOTG_SEG, q++)
        {
            DotTask tk ;
            const int32_t len = (w - s0 < DOTG_SEG) ? (w - s0) : DOTG_SEG ;
            tk.e = (int32_t) e ; tk.len = split ? -len : len ; tk.w0 = ws + s0 ;
            tasks [q] = tk ;
        }
    }
}

// Owner classes.  0: the owner fits one load of the cuckoo tables (or is dense): row-walk kernel,
// DOTG_CHUNK tasks per item.  1: a longer owner ("hub") whose index range fits DOTR_MAXPARTS bitmap
// parts: row-walk kernel with a shared-memory bitmap, big items.  2: any other hub: the segmented
// cuckoo kernel dotg_kernel<HUB>, big items.  3: a tiny owner (<= dotr_tiny_cap entries): the work item
// goes to a warp (dotr_warp_kernel).  flat == 0 (GB200_DOTR=0, for A/B runs) sends classes 0 and 3 to
// dotg_kernel<!HUB> and every hub to class 2.
struct DotgClasses { int64_t cap ; int64_t tiny ; int64_t bm_bits ; int flat ; int64_t chunk [4] ; } ;

__device__ __forceinline__ int dotg_class_of (const DMat &O, const DMat &M, int orient, int64_t v,
    const DotgClasses &K)
{
    int64_t ko = v ;
    if (!orient) ko = dm_vecpos (O, dm_vecname (M, v)) ;
    if (ko < 0) return 0 ;
    const int64_t o0 = O.p [ko], o1 = O.p [ko+1] ;
    const int64_t olen = o1 - o0 ;
    if (olen <= K.tiny && olen != O.vlen) return 3 ;     // tiny owner: a warp takes the item
    if (olen <= K.cap || olen == O.vlen) return 0 ;
    if (!K.flat) return 2 ;
    const int64_t lo0 = ((int64_t) __ldg (O.i + o0)) & ~(int64_t) 31 ;
    const int64_t nparts = (((int64_t) __ldg (O.i + o1 - 1)) - lo0) / K.bm_bits + 1 ;
    return (nparts <= DOTR_MAXPARTS) ? 1 : 2 ;
}

// All classes in ONE pass over the owners: an owner's task run is cut into the work items of its class,
// which are appended to that class's list with one atomic (the order of the items inside a list is
// immaterial: CTAs pull them from a counter).  Replaces a count pass, a scan over all owners, a host
// read-back and a fill pass PER CLASS -- at 4.2 M owners that was 0.6 ms of a multiply, and most of the
// set-up of one rank of an 8-GPU run.
struct DotgItemLists { DotItem *items [4] ; unsigned long long *count ; } ;

__global__ void __launch_bounds__ (256)
dotg_items_all_kernel (DMat O, DMat M, int orient, DotgClasses K,
    const int64_t *__restrict__ start, int64_t n, DotgItemLists L)
{
    // A block takes 256 consecutive owners at a time and claims the room of their items with ONE atomic
    // per class, so that inside such a stretch the items keep the owners' order (the items of one owner
    // are contiguous, neighbouring owners are neighbours in the list).  Handing the items out in the
    // order the threads' (or warps') atomics arrived cost DRAM traffic and time in the walk: 36.1 GB and
    // 21.9 ms per step against 32.5 GB and 21.7 ms (profiles/r2/tri_s22_dram_*.csv).  Stretches without
    // tasks (most of them on one rank of an N-GPU run) and classes nobody holds cost two barriers.
    constexpr unsigned FULL = 0xffffffffu ;
    __shared__ int64_t s_ws [33] ;
    __shared__ int64_t s_base [4] ;
    __shared__ unsigned s_mask [2] ;
    const int lane = threadIdx.x & 31 ;
    const int64_t nblk = (n + blockDim.x - 1) / blockDim.x ;
    int par = 0 ;
    for (int64_t blk = blockIdx.x ; blk < nblk ; blk += gridDim.x, par ^= 1)
    {
        const int64_t v = blk * blockDim.x + threadIdx.x ;
        int64_t s0 = 0, s1 = 0 ;
        if (v < n) { s0 = start [v] ; s1 = start [v+1] ; }
        int cls = -1 ;
        int64_t ch = 1, nch = 0 ;
        if (s1 > s0)
        {
            cls = dotg_class_of (O, M, orient, v, K) ;
            if (cls == 3 && K.tiny == 0) cls = 0 ;
            if (cls == 1 && !K.flat) cls = 2 ;
            ch = K.chunk [cls] ;
            nch = (s1 - s0 + ch - 1) / ch ;
        }
        // which classes does the stretch hold (s_mask alternates between two words, so that a thread that
        // is already in the next stretch does not clear the word a slower one still reads)
        if (threadIdx.x == 0) s_mask [par] = 0u ;
        __syncthreads () ;
        const unsigned wm = __reduce_or_sync (FULL, (cls >= 0) ? (1u << cls) : 0u) ;
        if (lane == 0 && wm) atomicOr (&s_mask [par], wm) ;
        __syncthreads () ;
        const unsigned present = s_mask [par] ;
        if (!present) continue ;
        int64_t q = 0 ;
        for (int c = 0 ; c < 4 ; c++)
        {
            if (!((present >> c) & 1u)) continue ;
            int64_t total ;
            const int64_t off = block_excl_scan_i64 ((cls == c) ? nch : 0, s_ws, total) ;
            if (threadIdx.x == 0) s_base [c] = (int64_t) atomicAdd (L.count + c, (unsigned long long) total) ;
            __syncthreads () ;
            if (cls == c) q = s_base [c] + off ;
        }
        for (int64_t e0 = s0 ; e0 < s1 ; e0 += ch, q++)
        {
            DotItem it ;
            it.owner = (int32_t) v ; it.pad = 0 ; it.e0 = e0 ; it.e1 = (e0 + ch < s1) ? (e0 + ch) : s1 ;
            L.items [cls][q] = it ;
        }
    }
}

static int pick_group (double len)
{
    int G = 1 ;
    while (G < 32 && G * 2 <= len) G <<= 1 ;
    return G ;
}

gb200_status run_dot (gb200_result_s *R, const gb200_dmatrix_s *M, int mask_comp,
    const gb200_dmatrix_s *Ad, const gb200_dmatrix_s *Bd, const gb200_semiring &s)
{
    Ctx &c = ctx () ;
    if (M != nullptr && !mask_comp)
    {
        // pattern-only operands (one repeated value) need no value loads (cached on the handle)
        GB200_TRY (ensure_iso (const_cast<gb200_dmatrix_s *> (Ad))) ;
        GB200_TRY (ensure_iso (const_cast<gb200_dmatrix_s *> (Bd))) ;
    }
    const DMat &A = Ad->v ;
    const DMat &B = Bd->v ;
    const int64_t cvlen = A.vdim, cvdim = B.vdim ;
    R->info.method_used = GB200_METHOD_DOT ;
    R->info.type_code = s.z_code ;
    R->info.mask_applied = (M != nullptr) ? 1 : 0 ;         // GB_AxB_dot.c:315
    // GB_AxB_dot.c passes (Mask_comp ? NULL : M) to GB_AxB_alloc
    const bool C_is_hyper = (cvdim > 1) &&
        (Ad->is_hyper_flag || Bd->is_hyper_flag || (M != nullptr && !mask_comp && M->is_hyper_flag)) ;

    int acc_size = 0 ;
    (void) identity_bits (s.z_code, s.add_opcode, &acc_size) ;
    const int zsize = type_size (s.z_code) ;
    const int is_bool = (s.z_code == GB200_BOOL) ;

    DMat Mv = DMat () ; DevBuf Mp2, Mi2 ;
    if (M != nullptr) GB200_TRY (filter_mask (M, Mv, Mp2, Mi2)) ;

    DevBuf nmatch ;
    GB200_TRY (nmatch.alloc (8)) ;
    GB200_CUDA (cudaMemsetAsync (nmatch.ptr, 0, 8, c.stream)) ;

    DotArgs da ;
    memset (&da, 0, sizeof (da)) ;
    da.A = A ; da.B = B ; da.M = Mv ;
    da.mult_op = s.mult_opcode ; da.flip = s.flipxy ;
    da.nmatch = nmatch.as<unsigned long long> () ;

    const double avgA = (A.nvec > 0) ? (double) A.nnz / (double) A.nvec : 0 ;
    const double avgB = (B.nvec > 0) ? (double) B.nnz / (double) B.nvec : 0 ;

    DevBuf Ci, Cx, ccum ;
    int64_t cnz = 0 ;
    gb200_status st ;

    if (M != nullptr && !mask_comp)
    {
        // ---- C<M> = A'*B : pairs are the (true) entries of M ----------------------------------
        const int64_t mnz = Mv.nnz ;
        DevBuf mvec, vals, flags, pos ;
        GB200_TRY (mvec.alloc ((mnz > 0 ? mnz : 1) * sizeof (int32_t))) ;
        GB200_TRY (vals.alloc ((size_t) (mnz > 0 ? mnz : 1) * acc_size)) ;
        GB200_TRY (flags.alloc (mnz > 0 ? mnz : 1)) ;
        GB200_TRY (pos.alloc ((mnz + 1) * sizeof (int64_t))) ;
        GB200_CUDA (cudaMemsetAsync (flags.ptr, 0, flags.bytes, c.stream)) ;
        if (mnz >= (int64_t) INT32_MAX) { set_error ("mask with >= 2^31 entries") ; return GB200_NOT_SUPPORTED ; }
        if (mnz > 0)
        {
            int asz = 0 ;
            const uint64_t ident = identity_bits (s.z_code, s.add_opcode, &asz) ;
            GB200_TRY (fill_bits (vals.ptr, acc_size, ident, mnz)) ;
            expand_vec_kernel <<<grid_cap ((Mv.nvec + 255) / 256, 16), 256, 0, c.stream>>> (Mv.p, Mv.nvec,
                mnz, mvec.as<int32_t> ()) ;
            count_launch () ;
            // ---- split the pairs by owner (the longer vector); regroup the A-owned ones by i -------
            const int64_t anvec = A.nvec ;
            const char *iso_env = getenv ("GB200_DOTG_ISO") ;        // 0: take the general path anyway
            const bool iso = (A.iso && B.iso) && !(iso_env != nullptr && atoi (iso_env) == 0) ;
            const int64_t cap = dotg_cap (iso) ;
            // 0: walk the whole list of every pair (for A/B measurements)
            const char *trim_env = getenv ("GB200_DOTG_TRIM") ;
            // 1 (default): always searched.  2 saves the searches of one-row lists (-0.08 ms of set-up) but
            // the untrimmed rows touch more sectors: 36.3 GB of DRAM traffic per step against 28.3 GB
            int trim = (trim_env != nullptr) ? atoi (trim_env) : 1 ;
            if (trim < 0 || trim > 2) trim = 1 ;
            // 0: warp per task / lane per task (dotg_kernel) instead of the row walk (kernels_dotr.cuh)
            const char *flat_env = getenv ("GB200_DOTR") ;
            const bool flat = !(flat_env != nullptr && atoi (flat_env) == 0) ;
            DevBuf w0buf, lk, nt0, cntA, offA, curA, toff0, off0, slist, nsmall, tasks ;
            GB200_TRY (w0buf.alloc (mnz * sizeof (int64_t))) ;
            GB200_TRY (lk.alloc (mnz * sizeof (int32_t))) ;
            GB200_TRY (nt0.alloc (mnz * sizeof (int32_t))) ;
            GB200_TRY (cntA.alloc ((anvec > 0 ? anvec : 1) * sizeof (int64_t))) ;
            GB200_TRY (curA.alloc ((anvec > 0 ? anvec : 1) * sizeof (int64_t))) ;
            GB200_TRY (offA.alloc ((anvec + 1) * sizeof (int64_t))) ;
            GB200_TRY (toff0.alloc ((mnz + 1) * sizeof (int64_t))) ;
            GB200_TRY (off0.alloc ((Mv.nvec + 1) * sizeof (int64_t))) ;
            GB200_TRY (slist.alloc (mnz * sizeof (int32_t))) ;
            GB200_TRY (nsmall.alloc (8)) ;
            GB200_CUDA (cudaMemsetAsync (cntA.ptr, 0, cntA.bytes, c.stream)) ;
            GB200_CUDA (cudaMemsetAsync (curA.ptr, 0, curA.bytes, c.stream)) ;
            GB200_CUDA (cudaMemsetAsync (nsmall.ptr, 0, 8, c.stream)) ;
            DevBuf infoA, infoB ;
            const bool sameAB = (A.p == B.p && A.i == B.i && A.nvec == B.nvec) ;
            const bool wide = (A.nnz >= (1LL << 32) || B.nnz >= (1LL << 32)) ;
            const size_t recsz = wide ? sizeof (VecInfo32) : sizeof (VecInfo16) ;
            GB200_TRY (infoA.alloc ((size_t) (anvec > 0 ? anvec : 1) * recsz)) ;
            if (!sameAB) GB200_TRY (infoB.alloc ((size_t) (B.nvec > 0 ? B.nvec : 1) * recsz)) ;
            const void *infoBp = sameAB ? infoA.ptr : infoB.ptr ;
            const int vgA = grid_cap ((anvec + 255) / 256, 16), vgB = grid_cap ((B.nvec + 255) / 256, 16) ;
            const int cg = grid_cap ((mnz + 255) / 256, 16) ;
            if (wide)
            {
                vec_info_kernel<VecInfo32> <<<vgA, 256, 0, c.stream>>> (A, infoA.as<VecInfo32> ()) ;
                if (!sameAB) vec_info_kernel<VecInfo32> <<<vgB, 256, 0, c.stream>>> (B, infoB.as<VecInfo32> ()) ;
                dotg_classify_kernel<VecInfo32> <<<cg, 256, 0, c.stream>>> (A, B, Mv, infoA.as<VecInfo32> (),
                    (const VecInfo32 *) infoBp, mvec.as<int32_t> (), mnz, trim, w0buf.as<int64_t> (),
                    lk.as<int32_t> (), nt0.as<int32_t> (), cntA.as<unsigned long long> (), slist.as<int32_t> (),
                    nsmall.as<unsigned int> ()) ;
            }
            else
            {
                vec_info_kernel<VecInfo16> <<<vgA, 256, 0, c.stream>>> (A, infoA.as<VecInfo16> ()) ;
                if (!sameAB) vec_info_kernel<VecInfo16> <<<vgB, 256, 0, c.stream>>> (B, infoB.as<VecInfo16> ()) ;
                dotg_classify_kernel<VecInfo16> <<<cg, 256, 0, c.stream>>> (A, B, Mv, infoA.as<VecInfo16> (),
                    (const VecInfo16 *) infoBp, mvec.as<int32_t> (), mnz, trim, w0buf.as<int64_t> (),
                    lk.as<int32_t> (), nt0.as<int32_t> (), cntA.as<unsigned long long> (), slist.as<int32_t> (),
                    nsmall.as<unsigned int> ()) ;
            }
            count_launch (sameAB ? 2 : 3) ;
            // B-owned pairs keep the mask's order (a scan of their task counts); A-owned: counting sort by i
            GB200_TRY (scan_i32 (nt0.as<int32_t> (), toff0.as<int64_t> (), mnz)) ;
            GB200_TRY (scan_i64 (cntA.as<int64_t> (), offA.as<int64_t> (), anvec)) ;
            int64_t nt_b = 0, nt_a = 0, ns = 0 ;
            {
                // the three totals in one round trip
                int64_t *hp = (int64_t *) c.pinned ;
                GB200_CUDA (cudaMemcpyAsync (hp, toff0.as<int64_t> () + mnz, 8, cudaMemcpyDeviceToHost, c.stream)) ;
                GB200_CUDA (cudaMemcpyAsync (hp + 1, offA.as<int64_t> () + anvec, 8, cudaMemcpyDeviceToHost, c.stream)) ;
                GB200_CUDA (cudaMemcpyAsync (hp + 2, nsmall.ptr, 8, cudaMemcpyDeviceToHost, c.stream)) ;
                GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
                nt_b = hp [0] ; nt_a = hp [1] ; ns = hp [2] & 0xffffffffLL ;
            }
            GB200_TRY (tasks.alloc ((nt_b + nt_a + 1) * sizeof (DotTask))) ;
            dotg_scatter_kernel <<<grid_cap ((mnz + 255) / 256, 16), 256, 0, c.stream>>> (A, Mv, mnz,
                w0buf.as<int64_t> (), lk.as<int32_t> (), toff0.as<int64_t> (), offA.as<int64_t> (),
                curA.as<unsigned long long> (), nt_b, tasks.as<DotTask> ()) ;
            // first task of every owner B(:,j): the scan at the start of the mask's vector j
            dot_cum_list_kernel <<<grid_cap ((Mv.nvec + 256) / 256, 8), 256, 0, c.stream>>> (Mv.p,
                toff0.as<int64_t> (), Mv.nvec, off0.as<int64_t> ()) ;
            count_launch (2) ;
            DotGArgs ga ;
            memset (&ga, 0, sizeof (ga)) ;
            // one work-item counter per launch (two orientations x four classes), then the failure flag
            DevBuf next_item ;
            GB200_TRY (next_item.alloc (16 * sizeof (unsigned long long))) ;
            GB200_CUDA (cudaMemsetAsync (next_item.ptr, 0, 16 * sizeof (unsigned long long), c.stream)) ;
            ga.failed = (unsigned int *) (next_item.as<unsigned long long> () + 8) ;
            ga.A = A ; ga.B = B ; ga.M = Mv ;
            ga.vals = vals.ptr ; ga.flags = flags.as<uint8_t> () ;
            ga.nmatch = nmatch.as<unsigned long long> () ;
            ga.mult_op = s.mult_opcode ; ga.flip = s.flipxy ;
            // ---- the owners' task runs of both orientations are cut into work items, one list per
            // owner class (hubs get big items) -----------------------------------------------------
            DotgClasses K [2] ;
            DevBuf itembuf [2][4], icount ;
            DotgItemLists IL [2] ;
            GB200_TRY (icount.alloc (8 * sizeof (unsigned long long))) ;
            GB200_CUDA (cudaMemsetAsync (icount.ptr, 0, 8 * sizeof (unsigned long long), c.stream)) ;
            for (int orient = 0 ; orient < 2 ; orient++)
            {
                const int64_t ntasks = orient ? nt_a : nt_b ;   // tasks of this orientation
                if (ntasks == 0) continue ;
                // task range of owner v: [otoff [v], otoff [v+1]) of this orientation's tasks
                const int64_t *otoff = orient ? offA.as<int64_t> () : off0.as<int64_t> () ;
                const int64_t nown = orient ? anvec : Mv.nvec ;
                // Tasks per hub item: big items amortise the owner's table, but there must also be
                // several items per resident block or a few big hubs serialise the launch (one rank of
                // an 8-GPU run holds an eighth of the hubs)
                int64_t hub_chunk = ntasks / (2 * 6 * (int64_t) c.sm_count * 2) ;
                hub_chunk = ((hub_chunk + 511) / 512) * 512 ;
                if (hub_chunk < 2048) hub_chunk = 2048 ;
                if (hub_chunk > DOTG_HUB_TASKS) hub_chunk = DOTG_HUB_TASKS ;
                if (getenv ("GB200_DOTG_HUB_CHUNK")) hub_chunk = atoll (getenv ("GB200_DOTG_HUB_CHUNK")) ;
                if (hub_chunk < 1 || hub_chunk > DOTG_HUB_TASKS) hub_chunk = DOTG_HUB_TASKS ;
                int64_t reg_chunk = DOTG_CHUNK ;
                if (getenv ("GB200_DOTG_CHUNK")) reg_chunk = atoll (getenv ("GB200_DOTG_CHUNK")) ;
                if (reg_chunk < 1) reg_chunk = DOTG_CHUNK ;
                DotgClasses &Ko = K [orient] ;
                Ko.cap = cap ; Ko.bm_bits = dotr_bm_bits (iso) ; Ko.flat = flat ? 1 : 0 ;
                // 0: no warp items (tiny owners go to the block kernel), for A/B runs
                const char *tiny_env = getenv ("GB200_DOTR_TINY") ;
                Ko.tiny = (flat && !(tiny_env != nullptr && atoi (tiny_env) == 0)) ? dotr_tiny_cap (iso) : 0 ;
                // smaller bitmap parts (a multiple of 32 indices): lets a test reach several parts
                if (getenv ("GB200_DOTR_BM_BITS"))
                {
                    const int64_t bb = (atoll (getenv ("GB200_DOTR_BM_BITS")) / 32) * 32 ;
                    if (bb >= 32 && bb <= Ko.bm_bits) Ko.bm_bits = bb ;
                }
                Ko.chunk [0] = reg_chunk ; Ko.chunk [1] = hub_chunk ; Ko.chunk [2] = hub_chunk ;
                Ko.chunk [3] = 256 ;
                // capacities: a class cannot hold more items than ntasks / chunk + one per owner with tasks
                const int64_t owners_max = (nown < ntasks) ? nown : ntasks ;
                for (int q = 0 ; q < 4 ; q++)
                {
                    GB200_TRY (itembuf [orient][q].alloc ((size_t) (ntasks / Ko.chunk [q] + owners_max + 1) * sizeof (DotItem))) ;
                    IL [orient].items [q] = itembuf [orient][q].as<DotItem> () ;
                }
                IL [orient].count = icount.as<unsigned long long> () + 4 * orient ;
                dotg_items_all_kernel <<<grid_cap ((nown + 255) / 256, 8), 256, 0, c.stream>>> (
                    orient ? A : B, Mv, orient, Ko, otoff, nown, IL [orient]) ;
                count_launch () ;
            }
            int64_t nitems_of [2][4] ;
            GB200_CUDA (cudaMemcpyAsync (c.pinned, icount.ptr, 8 * sizeof (unsigned long long),
                cudaMemcpyDeviceToHost, c.stream)) ;
            GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
            for (int q = 0 ; q < 8 ; q++) nitems_of [q >> 2][q & 3] = (int64_t) ((unsigned long long *) c.pinned) [q] ;
            // ---- the semiring kernels: independent of each other (disjoint pairs; pieces of a split
            // pair meet through the monoid's atomic), so they go to the side streams and one kernel's
            // tail is filled by the next one's blocks.  Hubs first (the big items), tiny owners and the
            // table-free small pairs last.  GB200_DOT_STREAMS=0: one after another on the main stream.
            const char *str_env = getenv ("GB200_DOT_STREAMS") ;
            const bool streams = !(str_env != nullptr && atoi (str_env) == 0) ;
            if (streams) GB200_TRY (group_begin ()) ;
            int nlaunch = 0 ;
            static const int class_order [4] = { 2, 1, 0, 3 } ;
            for (int co = 0 ; co < 4 ; co++)
            {
                const int cls = class_order [co] ;
                for (int orient = 0 ; orient < 2 ; orient++)
                {
                    const int64_t ntasks = orient ? nt_a : nt_b ;
                    const int64_t nitems = (ntasks > 0) ? nitems_of [orient][cls] : 0 ;
                    if (nitems == 0) continue ;
                    ga.tasks = tasks.as<DotTask> () + (orient ? nt_b : 0) ; ga.orient = orient ;
                    ga.bm_bits = K [orient].bm_bits ;
                    ga.next_item = next_item.as<unsigned long long> () + 4 * orient + cls ;
                    ga.items = IL [orient].items [cls] ; ga.nitems = nitems ;
                    int fam, per_sm, threads ;
                    if (cls == 2) { fam = iso ? FAM_DOTG_HUB_ISO : FAM_DOTG_HUB ; per_sm = 2 ; threads = DOTG_THREADS ; }
                    else if (cls == 1) { fam = iso ? FAM_DOTR_BM_ISO : FAM_DOTR_BM ; per_sm = 1 ; threads = DOTR_BM_THREADS ; }
                    else if (cls == 3) { fam = iso ? FAM_DOTR_WARP_ISO : FAM_DOTR_WARP ; per_sm = iso ? 3 : 2 ; threads = DOTR_THREADS ; }
                    else if (flat) { fam = iso ? FAM_DOTR_ISO : FAM_DOTR ; per_sm = iso ? 3 : 2 ; threads = DOTR_THREADS ; }
                    else { fam = iso ? FAM_DOTG_ISO : FAM_DOTG ; per_sm = iso ? 3 : 2 ; threads = DOTG_THREADS ; }
                    if (streams) group_use (nlaunch++) ;
                    if (!launch_typed (s.xy_code, fam, s.z_code, s.add_opcode, s.mult_opcode, &ga,
                        grid_cap (nitems, per_sm), threads))
                    { if (streams) group_end () ; set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
                }
            }
            if (ns > 0)
            {
                // short owner, shorter walk: a group of 4 lanes per pair, no table
                da.mode = DOT_MASK ; da.mvec = mvec.as<int32_t> () ; da.plist = slist.as<int32_t> () ;
                da.npairs = ns ; da.vals = vals.ptr ; da.flags = flags.as<uint8_t> () ; da.G = 4 ;
                if (streams) group_use (nlaunch++) ;
                if (!launch_typed (s.xy_code, FAM_DOT, s.z_code, s.add_opcode, s.mult_opcode, &da,
                    grid_cap ((ns + 63) / 64, 16), 256))
                { if (streams) group_end () ; set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
            }
            // joined before any buffer of this multiply is released (the block cache is stream-ordered
            // on the main stream)
            if (streams) GB200_TRY (group_end ()) ;
            // an owner whose cuckoo tables could not be built (never seen): every pair again, table-free
            int64_t failed = 0 ;
            GB200_TRY (read_i64 (next_item.as<int64_t> () + 8, &failed)) ;
            if (failed != 0)
            {
                GB200_CUDA (cudaMemsetAsync (flags.ptr, 0, flags.bytes, c.stream)) ;
                GB200_TRY (fill_bits (vals.ptr, acc_size, ident, mnz)) ;
                GB200_CUDA (cudaMemsetAsync (nmatch.ptr, 0, 8, c.stream)) ;
                da.mode = DOT_MASK ; da.mvec = mvec.as<int32_t> () ; da.plist = nullptr ;
                da.npairs = mnz ; da.vals = vals.ptr ; da.flags = flags.as<uint8_t> () ; da.G = 8 ;
                if (!launch_typed (s.xy_code, FAM_DOT, s.z_code, s.add_opcode, s.mult_opcode, &da,
                    grid_cap ((mnz + 31) / 32, 16), 256))
                { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
            }
        }
        GB200_TRY (scan_u8 (flags.as<uint8_t> (), pos.as<int64_t> (), mnz)) ;
        GB200_TRY (read_i64 (pos.as<int64_t> () + mnz, &cnz)) ;
        GB200_TRY (ccum.alloc ((Mv.nvec + 1) * sizeof (int64_t))) ;
        dot_cum_list_kernel <<<grid_cap ((Mv.nvec + 256) / 256, 8), 256, 0, c.stream>>> (Mv.p,
            pos.as<int64_t> (), Mv.nvec, ccum.as<int64_t> ()) ;
        count_launch () ;
        GB200_TRY (Ci.alloc ((cnz > 0 ? cnz : 1) * sizeof (int32_t))) ;
        GB200_TRY (Cx.alloc ((size_t) (cnz > 0 ? cnz : 1) * zsize)) ;
        if (cnz > 0)
        {
            dot_gather_kernel <<<grid_cap ((mnz + 255) / 256, 16), 256, 0, c.stream>>> (
                flags.as<uint8_t> (), pos.as<int64_t> (), mnz, Mv.i, A, vals.ptr, acc_size, zsize,
                is_bool, Ci.as<int32_t> (), Cx.ptr) ;
            count_launch () ;
        }
        GB200_CUDA (cudaGetLastError ()) ;
        st = assemble (R, Mv.nvec, Mv.hyper ? Mv.h : nullptr, Mv.hyper != 0, ccum, Ci, Cx, cnz,
            C_is_hyper, cvlen, cvdim) ;
    }
    else
    {
        // ---- C<!M> = A'*B or C = A'*B : pairs are (vector of A) x (vector of B), by slabs of B ---
        const int64_t anvec = A.nvec, bnvec = B.nvec ;
        DevBuf mposB, cnt ;
        const bool comp = (M != nullptr) ;
        if (comp)
        {
            GB200_TRY (mposB.alloc ((bnvec > 0 ? bnvec : 1) * sizeof (int64_t))) ;
            if (bnvec > 0)
            {
                GB200_TRY (launch_mask_pos (B, Mv, mposB.as<int64_t> ())) ;
            }
        }
        GB200_TRY (cnt.alloc ((bnvec > 0 ? bnvec : 1) * sizeof (int64_t))) ;
        GB200_CUDA (cudaMemsetAsync (cnt.ptr, 0, cnt.bytes, c.stream)) ;
        const int64_t max_pairs = 1LL << 27 ;
        int64_t nb = (anvec > 0) ? (max_pairs / anvec) : bnvec ;
        if (nb < 1) nb = 1 ;
        if (nb > bnvec) nb = bnvec ;
        std::vector<DevBuf> chunk_i, chunk_x ;
        std::vector<int64_t> chunk_nz ;
        da.mode = comp ? DOT_COMP : DOT_NONE ;
        da.mposB = mposB.as<int64_t> () ;
        // the list that is walked is A(:,i) when B(:,j) is dense, else the shorter of the two
        const bool Bdense = (B.nvec > 0 && B.nnz == B.nvec * B.vlen) ;
        da.G = pick_group (Bdense ? avgA : ((avgA < avgB) ? avgA : avgB)) ;
        for (int64_t jb0 = 0 ; jb0 < bnvec && anvec > 0 ; jb0 += nb)
        {
            const int64_t jb1 = (jb0 + nb < bnvec) ? (jb0 + nb) : bnvec ;
            const int64_t npairs = anvec * (jb1 - jb0) ;
            DevBuf vals, flags, pos, ci, cx ;
            GB200_TRY (vals.alloc ((size_t) npairs * acc_size)) ;
            GB200_TRY (flags.alloc (npairs)) ;
            GB200_TRY (pos.alloc ((npairs + 1) * sizeof (int64_t))) ;
            da.jb0 = jb0 ; da.jb1 = jb1 ; da.npairs = npairs ;
            da.vals = vals.ptr ; da.flags = flags.as<uint8_t> () ;
            const int64_t gpb = 256 / da.G ;
            if (!launch_typed (s.xy_code, FAM_DOT, s.z_code, s.add_opcode, s.mult_opcode, &da,
                grid_cap ((npairs + gpb - 1) / gpb, 16), 256))
            { set_error ("no kernel for this semiring") ; return GB200_NOT_SUPPORTED ; }
            GB200_TRY (scan_u8 (flags.as<uint8_t> (), pos.as<int64_t> (), npairs)) ;
            int64_t nz = 0 ;
            GB200_TRY (read_i64 (pos.as<int64_t> () + npairs, &nz)) ;
            dot_cnt_rect_kernel <<<grid_cap ((jb1 - jb0 + 255) / 256, 8), 256, 0, c.stream>>> (
                pos.as<int64_t> (), anvec, jb1 - jb0, cnt.as<int64_t> () + jb0) ;
            count_launch () ;
            GB200_TRY (ci.alloc ((nz > 0 ? nz : 1) * sizeof (int32_t))) ;
            GB200_TRY (cx.alloc ((size_t) (nz > 0 ? nz : 1) * zsize)) ;
            if (nz > 0)
            {
                dot_gather_kernel <<<grid_cap ((npairs + 255) / 256, 16), 256, 0, c.stream>>> (
                    flags.as<uint8_t> (), pos.as<int64_t> (), npairs, nullptr, A, vals.ptr, acc_size,
                    zsize, is_bool, ci.as<int32_t> (), cx.ptr) ;
                count_launch () ;
            }
            GB200_CUDA (cudaGetLastError ()) ;
            chunk_i.push_back (std::move (ci)) ;
            chunk_x.push_back (std::move (cx)) ;
            chunk_nz.push_back (nz) ;
            cnz += nz ;
        }
        GB200_TRY (ccum.alloc ((bnvec + 1) * sizeof (int64_t))) ;
        GB200_TRY (scan_i64 (cnt.as<int64_t> (), ccum.as<int64_t> (), bnvec)) ;
        if (chunk_i.size () == 1)
        {
            Ci = std::move (chunk_i [0]) ; Cx = std::move (chunk_x [0]) ;
        }
        else
        {
            GB200_TRY (Ci.alloc ((cnz > 0 ? cnz : 1) * sizeof (int32_t))) ;
            GB200_TRY (Cx.alloc ((size_t) (cnz > 0 ? cnz : 1) * zsize)) ;
            int64_t off = 0 ;
            for (size_t q = 0 ; q < chunk_i.size () ; q++)
            {
                if (chunk_nz [q] > 0)
                {
                    GB200_CUDA (cudaMemcpyAsync (Ci.as<int32_t> () + off, chunk_i [q].ptr,
                        chunk_nz [q] * sizeof (int32_t), cudaMemcpyDeviceToDevice, c.stream)) ;
                    GB200_CUDA (cudaMemcpyAsync ((char *) Cx.ptr + off * zsize, chunk_x [q].ptr,
                        (size_t) chunk_nz [q] * zsize, cudaMemcpyDeviceToDevice, c.stream)) ;
                }
                off += chunk_nz [q] ;
            }
        }
        st = assemble (R, bnvec, B.hyper ? B.h : nullptr, B.hyper != 0, ccum, Ci, Cx, cnz,
            C_is_hyper, cvlen, cvdim) ;
    }
    if (st != GB200_SUCCESS) return st ;
    int64_t nm = 0 ;
    GB200_TRY (read_i64 (nmatch.as<int64_t> (), &nm)) ;
    R->info.flops = nm ;
    return GB200_SUCCESS ;
}

} // namespace gb200

// =============================================================================================
// C ABI (part 2): the multiply
// =============================================================================================
using namespace gb200 ;

extern "C" {
#pragma GCC visibility push(default)

gb200_status gb200_AxB_device (gb200_result *out, gb200_dmatrix M, int mask_comp, gb200_dmatrix A,
    gb200_dmatrix B, const gb200_semiring *semiring, int do_adotb, int method)
{
    // saxpy-vs-dot was decided by the caller (GB_AxB_meta.c:266-366); HEAP and GUSTAVSON requests
    // run the same GPU saxpy.  Only the mask-policy flags of the request are looked at.
    const int mask_policy = (method & GB200_MASK_KEEP) ? 1 : ((method & GB200_MASK_DROP) ? 2 : 0) ;
    if (out == NULL || A == NULL || B == NULL || semiring == NULL) return GB200_INVALID ;
    *out = NULL ;
    GB200_TRY (ensure_init ()) ;
    gb200_semiring s = *semiring ;
    GB200_TRY (gb200_semiring_canonical (&s)) ;
    // operands of another built-in type are cast to the multiply operator's input type first (the
    // reference's typecasting generic path, GB_AxB_Gustavson.c:274-415): a temporary view of the
    // operand with a cast copy of its values; pointers, indices and hyperlist are shared
    gb200_dmatrix_s castA, castB ;
    if (A->v.type_code != s.xy_code)
    {
        std::lock_guard<std::recursive_mutex> lock (ctx ().mu) ;
        castA.v = A->v ; castA.is_hyper_flag = A->is_hyper_flag ;
        GB200_TRY (cast_values (A->v.x, A->v.type_code, s.xy_code, A->v.nnz, castA.x)) ;
        castA.v.x = castA.x.ptr ; castA.v.type_code = s.xy_code ; castA.v.iso = 0 ;
    }
    if (B->v.type_code != s.xy_code)
    {
        std::lock_guard<std::recursive_mutex> lock (ctx ().mu) ;
        castB.v = B->v ; castB.is_hyper_flag = B->is_hyper_flag ;
        GB200_TRY (cast_values (B->v.x, B->v.type_code, s.xy_code, B->v.nnz, castB.x)) ;
        castB.v.x = castB.x.ptr ; castB.v.type_code = s.xy_code ; castB.v.iso = 0 ;
    }
    if (A->v.type_code != s.xy_code) A = &castA ;
    if (B->v.type_code != s.xy_code) B = &castB ;
    const int64_t cvlen = do_adotb ? A->v.vdim : A->v.vlen ;
    const int64_t cvdim = B->v.vdim ;
    if ((do_adotb && A->v.vlen != B->v.vlen) || (!do_adotb && A->v.vdim != B->v.vlen)
        || (M != NULL && (M->v.vlen != cvlen || M->v.vdim != cvdim)))
    {
        set_error ("gb200_AxB_device: dimension mismatch") ;
        return GB200_INVALID ;
    }
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    gb200_result_s *R = new (std::nothrow) gb200_result_s () ;
    if (R == NULL) return GB200_OUT_OF_MEMORY ;
    memset (&R->info, 0, sizeof (R->info)) ;
    c.kev_used = 0 ;
    int64_t pm0 = 0, pu0 = 0 ;
    dev_pool_stats (&pm0, &pu0) ;
    const auto tw0 = std::chrono::steady_clock::now () ;
    c.mask_policy = mask_policy ;
    c.method_request = method & 0xffff ;
    cudaEventRecord (c.ev0, c.stream) ;
    gb200_status st ;
    if (vec_shape (A, B) && (M == NULL || (M->v.vdim == 1 && M->v.nvec == 1 && !M->v.hyper)))
        st = do_adotb ? run_dotv (R, M, mask_comp, A, B, s) : run_saxpyv (R, M, mask_comp, A, B, s) ;
    else
        st = do_adotb ? run_dot (R, M, mask_comp, A, B, s) : run_saxpy (R, M, mask_comp, A, B, s) ;
    if (st == GB200_SUCCESS)
    {
        cudaEventRecord (c.ev1, c.stream) ;
        cudaError_t e = cudaStreamSynchronize (c.stream) ;
        if (e != cudaSuccess)
        {
            set_error ("multiply failed on the device: %s", cudaGetErrorString (e)) ;
            st = GB200_CUDA_ERROR ;
        }
        else
        {
            float ms = 0 ;
            cudaEventElapsedTime (&ms, c.ev0, c.ev1) ;
            R->info.device_ms = ms ;
            double kms = 0 ;
            for (int q = 0 ; q + 1 < c.kev_used ; q += 2)
            {
                float t = 0 ;
                if (cudaEventElapsedTime (&t, c.kev [q], c.kev [q+1]) == cudaSuccess) kms += t ;
            }
            R->info.kernel_ms = kms ;
            if (getenv ("GB200_TRACE") != nullptr)
            {
                int64_t pm1 = 0, pu1 = 0 ;
                dev_pool_stats (&pm1, &pu1) ;
                fprintf (stderr, "gb200_AxB_device: device %.3f ms (semiring kernels %.3f), host wall %.3f ms, "
                    "%lld cudaMalloc calls taking %.3f ms\n", (double) ms, kms,
                    std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now () - tw0).count (),
                    (long long) (pm1 - pm0), (pu1 - pu0) * 1e-3) ;
            }
        }
    }
    if (st != GB200_SUCCESS)
    {
        cudaStreamSynchronize (c.stream) ;
        cudaGetLastError () ;
        delete R ;
        return st ;
    }
    c.multiplies++ ;
    *out = R ;
    return GB200_SUCCESS ;
}

gb200_status gb200_AxB_host (gb200_result *out, const gb200_matrix *M, int mask_comp,
    const gb200_matrix *A, const gb200_matrix *B, const gb200_semiring *semiring, int do_adotb,
    int method)
{
    if (out == NULL || A == NULL || B == NULL || semiring == NULL) return GB200_INVALID ;
    *out = NULL ;
    // decline before any transfer if the semiring is outside the built-in space
    gb200_semiring s = *semiring ;
    GB200_TRY (gb200_semiring_canonical (&s)) ;
    if (A->type_code < GB200_BOOL || A->type_code > GB200_FP64 || B->type_code < GB200_BOOL
        || B->type_code > GB200_FP64)
    {
        set_error ("operand of a user-defined type") ;
        return GB200_NOT_SUPPORTED ;
    }
    gb200_dmatrix dM = NULL, dA = NULL, dB = NULL ;
    gb200_status st = GB200_SUCCESS ;
    const bool trace = (getenv ("GB200_TRACE") != NULL) ;       // wall time of the phases, to stderr
    auto now = [] () { return std::chrono::duration<double, std::milli> (
        std::chrono::steady_clock::now ().time_since_epoch ()).count () ; } ;
    const double t0 = trace ? now () : 0 ;
    // operands may be the same host object (C=A*A; the tricount mask C<L>=L*U' is its own operand):
    // each distinct object crosses PCIe once
    auto same = [] (const gb200_matrix *X, const gb200_matrix *Y)
    {
        return X != NULL && Y != NULL && ((X == Y) || (X->p == Y->p && X->i == Y->i && X->x == Y->x
            && X->h == Y->h && X->vlen == Y->vlen && X->vdim == Y->vdim && X->nvec == Y->nvec
            && X->type_code == Y->type_code)) ;
    } ;
    // every distinct operand is resident once: from the residency cache (engine_cache.cu: only if the
    // host enabled it), else uploaded now and freed after the multiply
    bool cA = false, cB = false, cM = false ;
    st = cache_acquire (&dA, A, &cA) ;
    if (st == GB200_SUCCESS) { if (same (A, B)) dB = dA ; else st = cache_acquire (&dB, B, &cB) ; }
    if (st == GB200_SUCCESS && M != NULL)
    {
        if (same (M, A)) dM = dA ; else if (same (M, B)) dM = dB ; else st = cache_acquire (&dM, M, &cM) ;
    }
    const double t1 = trace ? now () : 0 ;
    if (st == GB200_SUCCESS) st = gb200_AxB_device (out, dM, mask_comp, dA, dB, semiring, do_adotb, method) ;
    const double t2 = trace ? now () : 0 ;
    if (dM != NULL && dM != dA && dM != dB) { if (cM) cache_release (dM) ; else gb200_dmatrix_free (&dM) ; }
    if (dB != NULL && dB != dA) { if (cB) cache_release (dB) ; else gb200_dmatrix_free (&dB) ; }
    if (dA != NULL) { if (cA) cache_release (dA) ; else gb200_dmatrix_free (&dA) ; }
    if (trace)
        fprintf (stderr, "gb200_AxB_host: upload %.3f ms, multiply %.3f ms (device %.3f ms), release %.3f ms\n",
            t1 - t0, t2 - t1, (st == GB200_SUCCESS && *out != NULL) ? (*out)->info.device_ms : 0.0,
            now () - t2) ;
    return st ;
}

gb200_status gb200_flopcount_device (gb200_dmatrix M, gb200_dmatrix A, gb200_dmatrix B,
    int64_t *Bflops_out, int64_t *total)
{
    if (A == NULL || B == NULL || total == NULL) return GB200_INVALID ;
    GB200_TRY (ensure_init ()) ;
    if (A->v.vdim != B->v.vlen) { set_error ("flopcount: dimension mismatch") ; return GB200_INVALID ; }
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    DevBuf flops, cum ;
    GB200_TRY (flopcount (M ? &M->v : nullptr, A->v, B->v, flops, cum, total)) ;
    if (Bflops_out != NULL)
    {
        GB200_CUDA (cudaMemcpyAsync (Bflops_out, cum.ptr, (B->v.nvec + 1) * sizeof (int64_t),
            cudaMemcpyDeviceToHost, c.stream)) ;
        GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
    }
    return GB200_SUCCESS ;
}

#pragma GCC visibility pop
} // extern "C"
