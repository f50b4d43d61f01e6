// engine_transpose.cu -- C = (ctype) A' on the device (the step immediately BEFORE
// the multiply whenever the descriptor or the formats ask for a transposed operand or mask, reference
// Source/GB_AxB_meta.c:203,247,311,328-337,355 -> Source/GB_transpose.c:38-985).
//
// Reference behaviour restated (GB_transpose.c:470-985, the general case avlen > 1, avdim > 1, no
// operator): entry A(i,j) -- index i in vector j -- becomes C(j,i): index j in vector i.  C is
// avdim-by-avlen in the reference's vlen/vdim terms, its vectors hold ascending indices, the values are
// A's, cast to ctype by the rule of GB_cast_array (GB_transpose_bucket.c, GB_builder.c).  Two host
// methods give the same entries: a bucket sort (T not hypersparse) and a quicksort of the tuples
// (GB_builder: T hypersparse, only non-empty vectors listed); GB_to_hyper_conform then decides the
// final form (the caller of this file says which form it wants and conforms with the reference's rule).
//
// GPU: a transpose is a STABLE sort of the entries by their index i: A's vectors are stored by ascending
// name j and entries of equal i keep their order, so every vector of C comes out ascending without a
// second sort.  Least-significant-digit radix sort, 8 bits per pass, ceil (log2 (avlen) / 8) passes over
// (key = i, payload = position of the entry in A); per pass a digit histogram per 4096-entry tile, the
// library's single-pass scan over (digit, tile), and a scatter in which a warp ranks its keys with
// ballots (the lanes of equal digit are found with 8 votes; one shared-memory update per digit and
// round, no atomics, no bank-conflict serialisation on banded matrices whose tiles hold one digit).
// Every pass streams 20 bytes per entry; the last step gathers j (from a position -> vector table
// written once, coalesced) and the values through the sorted positions.
#include "engine.cuh"
#include "scan.cuh"
#include "semiring.cuh"

namespace gb200 {

constexpr int TR_THREADS = 256 ;
constexpr int TR_WARPS = TR_THREADS / 32 ;
constexpr int TR_ROUNDS = 16 ;                          // keys per thread
constexpr int TR_TILE = TR_THREADS * TR_ROUNDS ;        // keys per block and tile
constexpr int TR_WARP_KEYS = 32 * TR_ROUNDS ;           // a warp's contiguous share of the tile
constexpr unsigned TR_FULL = 0xffffffffu ;

static inline int tr_grid (int64_t n, int per_sm)
{
    int64_t g = (n + 255) / 256, cap = (int64_t) ctx ().sm_count * per_sm ;
    if (g > cap) g = cap ;
    if (g < 1) g = 1 ;
    return (int) g ;
}

// the lanes of the warp whose key is valid and has the same 8-bit digit as this lane's (0 for a lane
// without a key).  All 32 lanes must call.
__device__ __forceinline__ unsigned tr_peers (unsigned d, bool valid)
{
    unsigned m = __ballot_sync (TR_FULL, valid) ;
    #pragma unroll
    for (int b = 0 ; b < 8 ; b++)
    {
        const bool bit = (d >> b) & 1u ;
        const unsigned bal = __ballot_sync (TR_FULL, bit) ;
        m &= bit ? bal : ~bal ;
    }
    return valid ? m : 0u ;
}

// hist [d * ntiles + t] = keys of tile t whose digit is d
__global__ void __launch_bounds__ (TR_THREADS) tr_hist_kernel (const uint32_t *__restrict__ keys, int64_t n,
    int shift, int64_t ntiles, int32_t *__restrict__ hist)
{
    __shared__ uint32_t cnt [TR_WARPS][256] ;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5 ;
    for (int64_t tile = blockIdx.x ; tile < ntiles ; tile += gridDim.x)
    {
        for (int d = lane ; d < 256 ; d += 32) cnt [w][d] = 0 ;
        __syncwarp () ;
        const int64_t wbase = tile * TR_TILE + (int64_t) w * TR_WARP_KEYS ;
        for (int r = 0 ; r < TR_ROUNDS ; r++)
        {
            const int64_t e = wbase + r * 32 + lane ;
            const bool valid = (e < n) ;
            const unsigned d = valid ? ((__ldg (keys + e) >> shift) & 255u) : 255u ;
            const unsigned m = tr_peers (d, valid) ;
            if (valid && lane == __ffs (m) - 1) cnt [w][d] += (uint32_t) __popc (m) ;
            __syncwarp () ;
        }
        __syncthreads () ;
        {
            const int d = threadIdx.x ;
            uint32_t s = 0 ;
            #pragma unroll
            for (int k = 0 ; k < TR_WARPS ; k++) s += cnt [k][d] ;
            hist [(int64_t) d * ntiles + tile] = (int32_t) s ;
        }
        __syncthreads () ;
    }
}

// base = exclusive scan of hist: where the keys of (digit, tile) start in the output.  Inside a tile
// the keys of one digit keep their order: warp after warp, round after round, lane after lane.
// pin == nullptr: the payload is the position itself (first pass)
__global__ void __launch_bounds__ (TR_THREADS) tr_scatter_kernel (const uint32_t *__restrict__ kin,
    const uint32_t *__restrict__ pin, int64_t n, int shift, int64_t ntiles, const int64_t *__restrict__ base,
    uint32_t *__restrict__ kout, uint32_t *__restrict__ pout)
{
    __shared__ uint32_t cnt [TR_WARPS][256] ;
    __shared__ int64_t gbase [256] ;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5 ;
    const unsigned below = (1u << lane) - 1u ;
    for (int64_t tile = blockIdx.x ; tile < ntiles ; tile += gridDim.x)
    {
        for (int d = lane ; d < 256 ; d += 32) cnt [w][d] = 0 ;
        __syncwarp () ;
        const int64_t wbase = tile * TR_TILE + (int64_t) w * TR_WARP_KEYS ;
        uint32_t key [TR_ROUNDS], pay [TR_ROUNDS], off [TR_ROUNDS] ;
        #pragma unroll
        for (int r = 0 ; r < TR_ROUNDS ; r++)
        {
            const int64_t e = wbase + r * 32 + lane ;
            const bool valid = (e < n) ;
            key [r] = valid ? __ldg (kin + e) : 0xffffffffu ;
            pay [r] = valid ? ((pin != nullptr) ? __ldg (pin + e) : (uint32_t) e) : 0u ;
        }
        #pragma unroll
        for (int r = 0 ; r < TR_ROUNDS ; r++)
        {
            const bool valid = (wbase + r * 32 + lane < n) ;
            const unsigned d = (key [r] >> shift) & 255u ;
            const unsigned m = tr_peers (d, valid) ;
            const int leader = valid ? (__ffs (m) - 1) : lane ;
            uint32_t old = 0 ;
            if (valid && lane == leader) { old = cnt [w][d] ; cnt [w][d] = old + (uint32_t) __popc (m) ; }
            old = __shfl_sync (TR_FULL, old, leader) ;
            off [r] = old + (uint32_t) __popc (m & below) ;
            __syncwarp () ;
        }
        __syncthreads () ;
        {
            // digit d: the warps' counts become the warps' starts inside the tile's run of digit d
            const int d = threadIdx.x ;
            uint32_t run = 0 ;
            #pragma unroll
            for (int k = 0 ; k < TR_WARPS ; k++) { const uint32_t t = cnt [k][d] ; cnt [k][d] = run ; run += t ; }
            gbase [d] = __ldg (base + (int64_t) d * ntiles + tile) ;
        }
        __syncthreads () ;
        #pragma unroll
        for (int r = 0 ; r < TR_ROUNDS ; r++)
        {
            if (wbase + r * 32 + lane < n)
            {
                const unsigned d = (key [r] >> shift) & 255u ;
                const int64_t q = gbase [d] + cnt [w][d] + off [r] ;
                kout [q] = key [r] ;
                pout [q] = pay [r] ;
            }
        }
        __syncthreads () ;
    }
}

// The same scatter with the tile's keys first laid out in SORTED order in shared memory (digit after digit,
// inside a digit in their stable order) and then written out by consecutive threads: the keys of one digit
// leave as one contiguous run instead of as 4-byte stores issued rounds apart (measured on RMAT 22, 64 M
// keys: the direct scatter's two random-digit passes take 2.0 and 1.8 ms for 1 GB of traffic each).
__global__ void __launch_bounds__ (TR_THREADS) tr_scatter_staged_kernel (const uint32_t *__restrict__ kin,
    const uint32_t *__restrict__ pin, int64_t n, int shift, int64_t ntiles, const int64_t *__restrict__ base,
    uint32_t *__restrict__ kout, uint32_t *__restrict__ pout)
{
    __shared__ uint32_t cnt [TR_WARPS][256] ;
    __shared__ int64_t gbase [256] ;                // start of (digit, tile) in the output MINUS dstart [digit]
    __shared__ uint32_t dstart [256] ;              // start of the digit's run inside the sorted tile
    __shared__ uint32_t wsum [TR_WARPS] ;
    __shared__ uint32_t skey [TR_TILE], spay [TR_TILE] ;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5 ;
    const unsigned below = (1u << lane) - 1u ;
    for (int64_t tile = blockIdx.x ; tile < ntiles ; tile += gridDim.x)
    {
        for (int d = lane ; d < 256 ; d += 32) cnt [w][d] = 0 ;
        __syncwarp () ;
        const int64_t wbase = tile * TR_TILE + (int64_t) w * TR_WARP_KEYS ;
        uint32_t key [TR_ROUNDS], pay [TR_ROUNDS], off [TR_ROUNDS] ;
        #pragma unroll
        for (int r = 0 ; r < TR_ROUNDS ; r++)
        {
            const int64_t e = wbase + r * 32 + lane ;
            const bool valid = (e < n) ;
            key [r] = valid ? __ldg (kin + e) : 0xffffffffu ;
            pay [r] = valid ? ((pin != nullptr) ? __ldg (pin + e) : (uint32_t) e) : 0u ;
        }
        #pragma unroll
        for (int r = 0 ; r < TR_ROUNDS ; r++)
        {
            const bool valid = (wbase + r * 32 + lane < n) ;
            const unsigned d = (key [r] >> shift) & 255u ;
            const unsigned m = tr_peers (d, valid) ;
            const int leader = valid ? (__ffs (m) - 1) : lane ;
            uint32_t old = 0 ;
            if (valid && lane == leader) { old = cnt [w][d] ; cnt [w][d] = old + (uint32_t) __popc (m) ; }
            old = __shfl_sync (TR_FULL, old, leader) ;
            off [r] = old + (uint32_t) __popc (m & below) ;
            __syncwarp () ;
        }
        __syncthreads () ;
        {
            // digit d = threadIdx.x: the warps' counts become the warps' starts inside the digit's run; the
            // digits' totals are scanned over the block into the runs' starts inside the sorted tile
            const int d = threadIdx.x ;
            uint32_t run = 0 ;
            #pragma unroll
            for (int k = 0 ; k < TR_WARPS ; k++) { const uint32_t t = cnt [k][d] ; cnt [k][d] = run ; run += t ; }
            uint32_t incl = run ;
            #pragma unroll
            for (int o = 1 ; o < 32 ; o <<= 1)
            {
                const uint32_t y = __shfl_up_sync (TR_FULL, incl, o) ;
                if (lane >= o) incl += y ;
            }
            if (lane == 31) wsum [w] = incl ;
            __syncthreads () ;
            uint32_t before = 0 ;
            for (int k = 0 ; k < w ; k++) before += wsum [k] ;
            const uint32_t ds = before + incl - run ;
            dstart [d] = ds ;
            gbase [d] = __ldg (base + (int64_t) d * ntiles + tile) - (int64_t) ds ;
        }
        __syncthreads () ;
        #pragma unroll
        for (int r = 0 ; r < TR_ROUNDS ; r++)
        {
            if (wbase + r * 32 + lane < n)
            {
                const unsigned d = (key [r] >> shift) & 255u ;
                const uint32_t sidx = dstart [d] + cnt [w][d] + off [r] ;
                skey [sidx] = key [r] ;
                spay [sidx] = pay [r] ;
            }
        }
        __syncthreads () ;
        const int64_t left = n - tile * TR_TILE ;
        const int nvalid = (left < TR_TILE) ? (int) left : TR_TILE ;
        for (int sidx = threadIdx.x ; sidx < nvalid ; sidx += TR_THREADS)
        {
            const uint32_t k = skey [sidx] ;
            const int64_t q = gbase [(k >> shift) & 255u] + sidx ;
            kout [q] = k ;
            pout [q] = spay [sidx] ;
        }
        __syncthreads () ;
    }
}

// vecof [e] = name of the vector of A that holds entry e.  A warp takes 32 consecutive vectors: the
// short ones are written by their lane, the long ones by the whole warp.
__global__ void tr_vecof_kernel (DMat A, int32_t *__restrict__ vecof)
{
    const int lane = threadIdx.x & 31 ;
    const int64_t wid = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5 ;
    const int64_t nw = ((int64_t) gridDim.x * blockDim.x) >> 5 ;
    for (int64_t v0 = wid * 32 ; v0 < A.nvec ; v0 += nw * 32)
    {
        const int64_t v = v0 + lane ;
        int64_t e0 = 0, e1 = 0 ;
        int32_t j = 0 ;
        if (v < A.nvec) { e0 = __ldg (A.p + v) ; e1 = __ldg (A.p + v + 1) ; j = (int32_t) dm_vecname (A, v) ; }
        const bool is_long = (e1 - e0 > 32) ;
        if (!is_long) for (int64_t e = e0 ; e < e1 ; e++) vecof [e] = j ;
        unsigned todo = __ballot_sync (TR_FULL, is_long) ;
        while (todo)
        {
            const int src = __ffs (todo) - 1 ;
            todo &= todo - 1 ;
            const int64_t s0 = __shfl_sync (TR_FULL, e0, src), s1 = __shfl_sync (TR_FULL, e1, src) ;
            const int32_t sj = __shfl_sync (TR_FULL, j, src) ;
            for (int64_t e = s0 + lane ; e < s1 ; e += 32) vecof [e] = sj ;
        }
    }
}

gb200_status launch_vecof (const DMat &A, int32_t *vecof)
{
    if (A.nnz <= 0 || A.nvec <= 0) return GB200_SUCCESS ;
    if (A.nvec == 1 && !A.hyper)
    {
        // an n-by-1 vector: every entry belongs to vector 0 (one warp would walk all of them in the kernel)
        GB200_CUDA (cudaMemsetAsync (vecof, 0, (size_t) A.nnz * sizeof (int32_t), ctx ().stream)) ;
        return GB200_SUCCESS ;
    }
    tr_vecof_kernel <<<tr_grid ((A.nvec + 31) / 32 * 32, 16), 256, 0, ctx ().stream>>> (A, vecof) ;
    count_launch () ;
    GB200_CUDA (cudaGetLastError ()) ;
    return GB200_SUCCESS ;
}

// head [q] = 1 where a new vector of C starts in the sorted keys
__global__ void tr_heads_kernel (const uint32_t *__restrict__ keys, int64_t n, uint8_t *__restrict__ head)
{
    for (int64_t q = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; q < n ; q += (int64_t) gridDim.x * blockDim.x)
        head [q] = (q == 0 || __ldg (keys + q) != __ldg (keys + q - 1)) ? 1 : 0 ;
}

// the non-empty vectors of C: names [s] = their index, cum [s] = where they start; cum [nsrc] = n
__global__ void tr_vectors_kernel (const uint32_t *__restrict__ keys, const uint8_t *__restrict__ head,
    const int64_t *__restrict__ pos, int64_t n, int64_t *__restrict__ names, int64_t *__restrict__ cum)
{
    for (int64_t q = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; q <= n ; q += (int64_t) gridDim.x * blockDim.x)
    {
        if (q == n) { cum [pos [n]] = n ; continue ; }
        if (!head [q]) continue ;
        const int64_t s = pos [q] ;
        names [s] = (int64_t) keys [q] ;
        cum [s] = q ;
    }
}

// Ci [q] = vector of A of the q-th sorted entry, Cx [q] = its value (tsz bytes, uncast)
__global__ void tr_gather_kernel (const uint32_t *__restrict__ perm, int64_t n, const int32_t *__restrict__ vecof,
    const unsigned char *__restrict__ Ax, int tsz, int32_t *__restrict__ Ci, unsigned char *__restrict__ Cx)
{
    for (int64_t q = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; q < n ; q += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t e = (int64_t) __ldg (perm + q) ;
        Ci [q] = __ldg (vecof + e) ;
        if (Ax == nullptr) continue ;                       // all values equal: filled by tr_fill_kernel
        if (tsz == 8) ((uint64_t *) Cx) [q] = __ldg ((const uint64_t *) Ax + e) ;
        else if (tsz == 4) ((uint32_t *) Cx) [q] = __ldg ((const uint32_t *) Ax + e) ;
        else if (tsz == 2) ((uint16_t *) Cx) [q] = __ldg ((const uint16_t *) Ax + e) ;
        else Cx [q] = __ldg (Ax + e) ;
    }
}

// Cx [q] = the one value x0 [0] of a matrix whose stored values are all equal (a pattern)
__global__ void tr_fill_kernel (const unsigned char *__restrict__ x0, int tsz, int64_t n, unsigned char *__restrict__ Cx)
{
    for (int64_t q = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; q < n ; q += (int64_t) gridDim.x * blockDim.x)
    {
        if (tsz == 8) ((uint64_t *) Cx) [q] = *(const uint64_t *) x0 ;
        else if (tsz == 4) ((uint32_t *) Cx) [q] = *(const uint32_t *) x0 ;
        else if (tsz == 2) ((uint16_t *) Cx) [q] = *(const uint16_t *) x0 ;
        else Cx [q] = *x0 ;
    }
}

} // namespace gb200

using namespace gb200 ;

extern "C" {
#pragma GCC visibility push(default)

// GB_to_hyper_conform (reference Source/GB_to_hyper_conform.c:38-58 with the tests of GB_to_hyper_test.c
// and GB_to_nonhyper_test.c, single precision as there): the form a matrix of `vdim` vectors, `k` of them
// non-empty, is left in when it starts out hypersparse or not
static bool conformed_hyper (bool is_hyper, int64_t k, int64_t vdim, double hyper_ratio)
{
    const float n = (float) vdim, r = (float) hyper_ratio ;
    if (k < 0) k = 0 ;
    if ((float) k > n) k = (int64_t) n ;
    if (!is_hyper) return (n > 1 && ((float) k) <= n * r) ;
    return !(n <= 1 || ((float) k) > n * r * 2) ;
}

gb200_status gb200_transpose_device (gb200_result *out, gb200_dmatrix Ad, int ctype_code, int result_hyper,
    double hyper_ratio)
{
    if (out == NULL || Ad == NULL) return GB200_INVALID ;
    *out = NULL ;
    if (ctype_code < GB200_BOOL || ctype_code > GB200_FP64)
    {
        set_error ("transpose into a user-defined type") ;
        return GB200_NOT_SUPPORTED ;
    }
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    const DMat &A = Ad->v ;
    const int64_t n = A.nnz ;
    if (n >= (int64_t) 0xffffffffLL)
    {
        set_error ("transpose of %lld entries: positions are kept in 32 bits", (long long) n) ;
        return GB200_NOT_SUPPORTED ;
    }
    const int asz = type_size (A.type_code) ;
    gb200_result_s *R = new (std::nothrow) gb200_result_s () ;
    if (R == NULL) return GB200_OUT_OF_MEMORY ;
    memset (&R->info, 0, sizeof (R->info)) ;
    auto body = [&] () -> gb200_status
    {
        cudaEventRecord (c.ev0, c.stream) ;
        int bits = 0 ;
        while (bits < 32 && ((int64_t) 1 << bits) < A.vlen) bits++ ;
        const int passes = (bits + 7) / 8 ;                 // 0: every entry has index 0
        // A/B switches (tools/transpose_bench.py --ab): GB200_TR_STAGED=0 the direct scatter
        const char *env_staged = getenv ("GB200_TR_STAGED") ;
        const bool staged = (env_staged == nullptr || atoi (env_staged) != 0) ;
        const int64_t ntiles = (n + TR_TILE - 1) / TR_TILE ;
        const int grid = (int) ((ntiles < (int64_t) c.sm_count * 4) ? (ntiles > 0 ? ntiles : 1)
            : (int64_t) c.sm_count * 4) ;
        DevBuf kbuf [2], pbuf [2], hist, base, vecof ;
        const uint32_t *keys = (const uint32_t *) A.i ;     // indices are non-negative 32-bit
        const uint32_t *perm = nullptr ;                    // nullptr: the identity
        GB200_TRY (vecof.alloc ((size_t) (n > 0 ? n : 1) * sizeof (int32_t))) ;
        if (n > 0)
        {
            GB200_TRY (launch_vecof (A, vecof.as<int32_t> ())) ;
            if (passes > 0)
            {
                for (int k = 0 ; k < 2 ; k++)
                {
                    GB200_TRY (kbuf [k].alloc ((size_t) n * sizeof (uint32_t))) ;
                    GB200_TRY (pbuf [k].alloc ((size_t) n * sizeof (uint32_t))) ;
                }
                GB200_TRY (hist.alloc ((size_t) ntiles * 256 * sizeof (int32_t))) ;
                GB200_TRY (base.alloc (((size_t) ntiles * 256 + 1) * sizeof (int64_t))) ;
            }
            for (int pass = 0 ; pass < passes ; pass++)
            {
                const int shift = 8 * pass ;
                uint32_t *kout = kbuf [pass & 1].as<uint32_t> (), *pout = pbuf [pass & 1].as<uint32_t> () ;
                tr_hist_kernel <<<grid, TR_THREADS, 0, c.stream>>> (keys, n, shift, ntiles, hist.as<int32_t> ()) ;
                count_launch () ;
                GB200_TRY (scan_i32 (hist.as<int32_t> (), base.as<int64_t> (), ntiles * 256)) ;
                if (staged)
                    tr_scatter_staged_kernel <<<grid, TR_THREADS, 0, c.stream>>> (keys, perm, n, shift, ntiles,
                        base.as<int64_t> (), kout, pout) ;
                else
                    tr_scatter_kernel <<<grid, TR_THREADS, 0, c.stream>>> (keys, perm, n, shift, ntiles,
                        base.as<int64_t> (), kout, pout) ;
                count_launch () ;
                keys = kout ; perm = pout ;
            }
        }
        GB200_CUDA (cudaGetLastError ()) ;
        // the vectors of C
        DevBuf head, pos, names, cum, Ci, Craw, Cx ;
        GB200_TRY (head.alloc (n > 0 ? n : 1)) ;
        GB200_TRY (pos.alloc ((n + 1) * sizeof (int64_t))) ;
        if (n > 0)
        {
            tr_heads_kernel <<<tr_grid (n, 16), 256, 0, c.stream>>> (keys, n, head.as<uint8_t> ()) ;
            count_launch () ;
        }
        GB200_TRY (scan_u8 (head.as<uint8_t> (), pos.as<int64_t> (), n)) ;
        int64_t nsrc = 0 ;
        GB200_TRY (read_i64 (pos.as<int64_t> () + n, &nsrc)) ;
        GB200_TRY (names.alloc ((nsrc > 0 ? nsrc : 1) * sizeof (int64_t))) ;
        GB200_TRY (cum.alloc ((nsrc + 1) * sizeof (int64_t))) ;
        tr_vectors_kernel <<<tr_grid (n + 1, 16), 256, 0, c.stream>>> (keys, head.as<uint8_t> (),
            pos.as<int64_t> (), n, names.as<int64_t> (), cum.as<int64_t> ()) ;
        count_launch () ;
        GB200_TRY (Ci.alloc ((size_t) (n > 0 ? n : 1) * sizeof (int32_t))) ;
        GB200_TRY (Craw.alloc ((size_t) (n > 0 ? n : 1) * asz)) ;
        if (n > 0)
        {
            if (perm != nullptr)
            {
                // a pattern (all stored values equal, e.g. the adjacency matrix of a graph) needs no gather of
                // its values: half of the random reads of this step
                const char *env_iso = getenv ("GB200_TR_ISO") ;
                bool iso = false ;
                if (env_iso == nullptr || atoi (env_iso) != 0)
                {
                    GB200_TRY (ensure_iso (Ad)) ;
                    iso = (Ad->v.iso != 0) ;
                }
                tr_gather_kernel <<<tr_grid (n, 16), 256, 0, c.stream>>> (perm, n, vecof.as<int32_t> (),
                    iso ? nullptr : (const unsigned char *) A.x, asz, Ci.as<int32_t> (), (unsigned char *) Craw.ptr) ;
                count_launch () ;
                if (iso)
                {
                    tr_fill_kernel <<<tr_grid (n, 16), 256, 0, c.stream>>> ((const unsigned char *) A.x, asz, n,
                        (unsigned char *) Craw.ptr) ;
                    count_launch () ;
                }
            }
            else
            {
                // a single vector of C: the entries stay where they are
                GB200_CUDA (cudaMemcpyAsync (Ci.ptr, vecof.ptr, (size_t) n * sizeof (int32_t),
                    cudaMemcpyDeviceToDevice, c.stream)) ;
                GB200_CUDA (cudaMemcpyAsync (Craw.ptr, A.x, (size_t) n * asz, cudaMemcpyDeviceToDevice, c.stream)) ;
            }
        }
        if (ctype_code != A.type_code) GB200_TRY (cast_values (Craw.ptr, A.type_code, ctype_code, n, Cx)) ;
        else Cx = std::move (Craw) ;
        GB200_CUDA (cudaGetLastError ()) ;
        R->info.type_code = ctype_code ;
        R->info.method_used = 0 ; R->info.mask_applied = 0 ; R->info.flops = n ;
        bool c_hyper = (result_hyper != 0) ;
        if (hyper_ratio >= 0) c_hyper = conformed_hyper (c_hyper, nsrc, A.vlen, hyper_ratio) ;
        GB200_TRY (assemble (R, nsrc, names.as<int64_t> (), true, cum, Ci, Cx, n, c_hyper, A.vdim, A.vlen)) ;
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

gb200_status gb200_transpose_host (gb200_result *out, const gb200_matrix *A, int ctype_code, int result_hyper,
    double hyper_ratio)
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
    gb200_status st = gb200_transpose_device (out, dA, ctype_code, result_hyper, hyper_ratio) ;
    if (cached) cache_release (dA) ; else gb200_dmatrix_free (&dA) ;
    return st ;
}

#pragma GCC visibility pop
} // extern "C"
