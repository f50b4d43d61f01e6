// engine_util.cu -- context, upload / fetch, scan wrappers, mask filtering, result assembly and
// the parts of the C ABI that do not depend on the multiply itself.
#include <cstdarg>
#include <cstdlib>
#include <cmath>
#include <limits>
#include <type_traits>
#include <chrono>
#include "engine.cuh"
#include "scan.cuh"
#include "semiring.cuh"

namespace gb200 {

static thread_local char g_err [512] = "" ;

void set_error (const char *fmt, ...)
{
    va_list ap ; va_start (ap, fmt) ;
    vsnprintf (g_err, sizeof (g_err), fmt, ap) ;
    va_end (ap) ;
    if (getenv ("GB200_VERBOSE")) fprintf (stderr, "[gb_b200] error: %s\n", g_err) ;
}

Ctx &ctx () { static Ctx c ; return c ; }

static gb200_status do_init (int device)
{
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    if (c.ready) return GB200_SUCCESS ;
    int n = 0 ;
    cudaError_t e = cudaGetDeviceCount (&n) ;
    if (e != cudaSuccess || n == 0)
    {
        cudaGetLastError () ;
        set_error ("no usable CUDA device (%s); this library has no CPU path",
            (e != cudaSuccess) ? cudaGetErrorString (e) : "device count is 0") ;
        return GB200_NO_DEVICE ;
    }
    if (device < 0)
    {
        const char *env = getenv ("GB200_DEVICE") ;
        device = env ? atoi (env) : 0 ;
    }
    if (device >= n) { set_error ("device %d requested, %d present", device, n) ; return GB200_NO_DEVICE ; }
    GB200_CUDA (cudaSetDevice (device)) ;
    cudaDeviceProp prop ;
    GB200_CUDA (cudaGetDeviceProperties (&prop, device)) ;
    if (prop.major < 10)
    {
        set_error ("device %d is sm_%d%d; this library is built for sm_100a only", device,
            prop.major, prop.minor) ;
        return GB200_NO_DEVICE ;
    }
    c.device = device ;
    c.sm_count = prop.multiProcessorCount ;
    GB200_CUDA (cudaStreamCreateWithFlags (&c.stream, cudaStreamNonBlocking)) ;
    GB200_CUDA (cudaEventCreate (&c.ev0)) ;
    GB200_CUDA (cudaEventCreate (&c.ev1)) ;
    GB200_CUDA (cudaEventCreateWithFlags (&c.fork_ev, cudaEventDisableTiming)) ;
    for (int k = 0 ; k < Ctx::NSIDE ; k++)
    {
        GB200_CUDA (cudaStreamCreateWithFlags (&c.side [k], cudaStreamNonBlocking)) ;
        GB200_CUDA (cudaEventCreateWithFlags (&c.side_done [k], cudaEventDisableTiming)) ;
    }
    c.pinned_bytes = 1 << 16 ;
    GB200_CUDA (cudaMallocHost (&c.pinned, c.pinned_bytes)) ;
    c.ready = true ;
    return GB200_SUCCESS ;
}

// ---------------------------------------------------------------------------------------------
// device block cache (see DevBuf in engine.cuh)
// ---------------------------------------------------------------------------------------------
struct DevPool
{
    std::mutex mu ;
    std::multimap<size_t, void *> cache ;       // free blocks by capacity
    size_t cached = 0 ;
} ;
static DevPool &dev_pool () { static DevPool *p = new DevPool () ; return *p ; }   // never destroyed

static size_t dev_size_class (size_t n)
{
    size_t c = 512 ;
    while (c < n) c <<= 1 ;
    if (c <= (1u << 20) || c == n) return c ;
    // eight classes per octave above 1 MiB: at most 12.5 % of slack
    const size_t lo = c >> 1, step = lo >> 3 ;
    for (size_t q = lo + step ; q < c ; q += step) if (q >= n) return q ;
    return c ;
}

// how often the pool had to go to the driver, and how long that took (GB200_TRACE prints the figures
// of every multiply: a cudaMalloc inside a multiply stalls the stream it is queued on)
static std::atomic<int64_t> g_pool_mallocs {0}, g_pool_malloc_us {0} ;
void dev_pool_stats (int64_t *mallocs, int64_t *malloc_us)
{
    *mallocs = g_pool_mallocs.load () ; *malloc_us = g_pool_malloc_us.load () ;
}

void *dev_pool_alloc (size_t nbytes, size_t *capacity)
{
    DevPool &dp = dev_pool () ;
    const size_t cap = dev_size_class (nbytes) ;
    {
        std::lock_guard<std::mutex> lock (dp.mu) ;
        // the smallest cached block that fits, if it is at most twice the size asked for: an iterative
        // caller whose operands shrink from call to call (k-truss, BFS levels) asks for slightly different
        // sizes every time, and a cudaMalloc of a large block inside a multiply stalls it for tens of
        // milliseconds (measured: 8 calls, 318 ms, in a 45 ms multiply)
        auto it = dp.cache.lower_bound (cap) ;
        if (it != dp.cache.end () && it->first <= 2 * cap)
        {
            void *p = it->second ;
            const size_t have = it->first ;
            dp.cache.erase (it) ;
            dp.cached -= have ;
            *capacity = have ;
            return p ;
        }
    }
    void *p = nullptr ;
    const auto t0 = std::chrono::steady_clock::now () ;
    cudaError_t e = cudaMalloc (&p, cap) ;
    if (e != cudaSuccess)
    {
        cudaGetLastError () ;
        dev_pool_trim () ;                      // drop the cache and try once more
        e = cudaMalloc (&p, cap) ;
        if (e != cudaSuccess) { cudaGetLastError () ; return nullptr ; }
    }
    g_pool_mallocs++ ;
    g_pool_malloc_us += (int64_t) std::chrono::duration<double, std::micro> (std::chrono::steady_clock::now () - t0).count () ;
    *capacity = cap ;
    return p ;
}

// Upper bound of the free-block cache: GB200_DEVICE_CACHE_MB, else half of the device's memory, so
// that other users of the same GPU (torch, NCCL) are not starved by blocks this library holds free.
static size_t dev_cache_limit ()
{
    static size_t limit = 0 ;
    if (limit == 0)
    {
        const char *env = getenv ("GB200_DEVICE_CACHE_MB") ;
        size_t free_b = 0, total_b = 0 ;
        if (env != nullptr && atoll (env) >= 0) limit = ((size_t) atoll (env) << 20) + 1 ;
        else if (cudaMemGetInfo (&free_b, &total_b) == cudaSuccess && total_b > 0) limit = total_b / 2 + 1 ;
        else { cudaGetLastError () ; limit = ((size_t) 64 << 30) + 1 ; }
    }
    return limit ;
}

void dev_pool_free (void *ptr, size_t capacity)
{
    if (ptr == nullptr) return ;
    DevPool &dp = dev_pool () ;
    std::vector<void *> drop ;
    {
        std::lock_guard<std::mutex> lock (dp.mu) ;
        dp.cache.emplace (capacity, ptr) ;
        dp.cached += capacity ;
        const size_t limit = dev_cache_limit () ;
        while (dp.cached > limit && !dp.cache.empty ())     // largest first
        {
            auto it = std::prev (dp.cache.end ()) ;
            dp.cached -= it->first ;
            drop.push_back (it->second) ;
            dp.cache.erase (it) ;
        }
    }
    if (drop.empty ()) return ;
    // the blocks may still be in use by queued work of the library's stream
    Ctx &c = ctx () ;
    if (c.stream != nullptr) cudaStreamSynchronize (c.stream) ;
    for (void *q : drop) cudaFree (q) ;
    cudaGetLastError () ;
}

void dev_pool_trim ()
{
    DevPool &dp = dev_pool () ;
    std::multimap<size_t, void *> drop ;
    {
        std::lock_guard<std::mutex> lock (dp.mu) ;
        drop.swap (dp.cache) ;
        dp.cached = 0 ;
    }
    if (drop.empty ()) return ;
    // blocks may still be in use by queued work of the library's stream
    Ctx &c = ctx () ;
    if (c.stream != nullptr) cudaStreamSynchronize (c.stream) ;
    for (auto &kv : drop) cudaFree (kv.second) ;
    cudaGetLastError () ;
}

gb200_status ensure_init ()
{
    Ctx &c = ctx () ;
    bool ready ;
    { std::lock_guard<std::recursive_mutex> lock (c.mu) ; ready = c.ready ; }
    if (!ready) { GB200_TRY (do_init (-1)) ; }
    GB200_CUDA (cudaSetDevice (c.device)) ;
    return GB200_SUCCESS ;
}

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
__global__ void narrow_idx_kernel (const int64_t *__restrict__ in, int32_t *__restrict__ out, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < n ;
        t += (int64_t) gridDim.x * blockDim.x) out [t] = (int32_t) in [t] ;
}

__global__ void widen_idx_kernel (const int32_t *__restrict__ in, int64_t *__restrict__ out, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < n ;
        t += (int64_t) gridDim.x * blockDim.x) out [t] = (int64_t) in [t] ;
}

template <class W>
__global__ void fill_kernel (W *__restrict__ dst, W v, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < n ;
        t += (int64_t) gridDim.x * blockDim.x) dst [t] = v ;
}

static inline int grid_for (int64_t n, int block = 256, int per_sm = 8)
{
    int64_t g = (n + block - 1) / block ;
    int64_t cap = (int64_t) ctx ().sm_count * per_sm ;
    if (g > cap) g = cap ;
    if (g < 1) g = 1 ;
    return (int) g ;
}

gb200_status fill_bits (void *dst, int elem_size, uint64_t bits, int64_t n)
{
    if (n <= 0) return GB200_SUCCESS ;
    cudaStream_t st = ctx ().stream ;
    if (elem_size == 4) fill_kernel<uint32_t> <<<grid_for (n), 256, 0, st>>> ((uint32_t *) dst, (uint32_t) bits, n) ;
    else if (elem_size == 8) fill_kernel<uint64_t> <<<grid_for (n), 256, 0, st>>> ((uint64_t *) dst, bits, n) ;
    else if (elem_size == 1) fill_kernel<uint8_t> <<<grid_for (n), 256, 0, st>>> ((uint8_t *) dst, (uint8_t) bits, n) ;
    else if (elem_size == 2) fill_kernel<uint16_t> <<<grid_for (n), 256, 0, st>>> ((uint16_t *) dst, (uint16_t) bits, n) ;
    else { set_error ("fill_bits: bad element size %d", elem_size) ; return GB200_INVALID ; }
    count_launch () ;
    GB200_CUDA (cudaGetLastError ()) ;
    return GB200_SUCCESS ;
}

// identity of the monoid in accumulator representation (see AccOf in semiring.cuh)
template <class Z> static uint64_t ident_bits_t (int add)
{
    using acc_t = typename AccOf<Z>::type ;
    acc_t v ;
    switch (add)
    {
        case GB200_MIN   : v = Monoid<Z, GB200_MIN>::identity () ; break ;
        case GB200_MAX   : v = Monoid<Z, GB200_MAX>::identity () ; break ;
        case GB200_PLUS  : v = (acc_t) 0 ; break ;
        case GB200_TIMES : v = (acc_t) 1 ; break ;
        case GB200_LOR   : v = (acc_t) 0 ; break ;
        case GB200_LAND  : v = (acc_t) 1 ; break ;
        case GB200_LXOR  : v = (acc_t) 0 ; break ;
        default          : v = (acc_t) 1 ; break ;      // EQ
    }
    uint64_t bits = 0 ;
    memcpy (&bits, &v, sizeof (acc_t)) ;
    return bits ;
}
template <> uint64_t ident_bits_t<bool> (int add)
{
    return (add == GB200_LAND || add == GB200_EQ) ? 1u : 0u ;
}

uint64_t identity_bits (int z_code, int add, int *acc_size)
{
    *acc_size = (type_size (z_code) < 4) ? 4 : type_size (z_code) ;
    switch (z_code)
    {
        case GB200_BOOL   : return ident_bits_t<bool> (add) ;
        case GB200_INT8   : return ident_bits_t<int8_t> (add) ;
        case GB200_UINT8  : return ident_bits_t<uint8_t> (add) ;
        case GB200_INT16  : return ident_bits_t<int16_t> (add) ;
        case GB200_UINT16 : return ident_bits_t<uint16_t> (add) ;
        case GB200_INT32  : return ident_bits_t<int32_t> (add) ;
        case GB200_UINT32 : return ident_bits_t<uint32_t> (add) ;
        case GB200_INT64  : return ident_bits_t<int64_t> (add) ;
        case GB200_UINT64 : return ident_bits_t<uint64_t> (add) ;
        case GB200_FP32   : return ident_bits_t<float> (add) ;
        default           : return ident_bits_t<double> (add) ;
    }
}

// ---------------------------------------------------------------------------------------------
// scans
// ---------------------------------------------------------------------------------------------
template <class InT>
static gb200_status scan_any (const InT *in, int64_t *out, int64_t n)
{
    Ctx &c = ctx () ;
    const int64_t ntiles = (n + 1 + SCAN_TILE - 1) / SCAN_TILE ;
    DevBuf state ;
    // layout: [ticket (16 B)] [flag int x ntiles] [aggregate] [inclusive]
    const size_t off_flag = 16 ;
    const size_t off_agg = off_flag + ((ntiles * sizeof (int) + 15) / 16) * 16 ;
    const size_t off_inc = off_agg + ntiles * sizeof (int64_t) ;
    const size_t total = off_inc + ntiles * sizeof (int64_t) ;
    GB200_TRY (state.alloc (total)) ;
    GB200_CUDA (cudaMemsetAsync (state.ptr, 0, off_agg, c.stream)) ;
    ScanState st ;
    st.ticket = (unsigned int *) state.ptr ;
    st.flag = (int *) ((char *) state.ptr + off_flag) ;
    st.aggregate = (int64_t *) ((char *) state.ptr + off_agg) ;
    st.inclusive = (int64_t *) ((char *) state.ptr + off_inc) ;
    scan_kernel<InT> <<<(unsigned) ntiles, SCAN_THREADS, 0, c.stream>>> (in, out, n, st) ;
    count_launch () ;
    GB200_CUDA (cudaGetLastError ()) ;
    return GB200_SUCCESS ;
}

gb200_status scan_i64 (const int64_t *in, int64_t *out, int64_t n) { return scan_any<int64_t> (in, out, n) ; }
gb200_status scan_u8  (const uint8_t *in, int64_t *out, int64_t n) { return scan_any<uint8_t> (in, out, n) ; }
gb200_status scan_i32 (const int32_t *in, int64_t *out, int64_t n) { return scan_any<int32_t> (in, out, n) ; }

gb200_status read_i64 (const int64_t *dptr, int64_t *host)
{
    Ctx &c = ctx () ;
    GB200_CUDA (cudaMemcpyAsync (c.pinned, dptr, sizeof (int64_t), cudaMemcpyDeviceToHost, c.stream)) ;
    GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
    *host = *(int64_t *) c.pinned ;
    return GB200_SUCCESS ;
}

// ---------------------------------------------------------------------------------------------
// typecasting of built-in operand types.  The reference's generic path casts every entry of A and B
// to the multiply operator's input type as it is used (Source/GB_AxB_Gustavson.c:360-404,
// GB_AxB_dot.c:189-305, GB_cast_factory); casting is a pure function of the entry, so the value
// array is cast once up front and the built-in worker runs on the copy.  Rule: GB_CAST,
// Source/GB.h:2925-2947 -- float/double NaN -> integer 0, +Inf / -Inf -> the largest / smallest
// integer, to bool: x != 0, anything else the C cast.  (A finite value outside the integer's range is
// undefined behaviour in C, Source/GB.h:2902-2913; the device saturates.)
// ---------------------------------------------------------------------------------------------
template <class To, class From> __device__ __forceinline__ To cast_one (From x)
{
    if constexpr (std::is_same<To, bool>::value) return (x != (From) 0) ;
    else if constexpr (std::is_integral<To>::value && std::is_floating_point<From>::value)
    {
        if (isnan (x)) return (To) 0 ;
        if (isinf (x)) return (x > 0) ? std::numeric_limits<To>::max () : std::numeric_limits<To>::min () ;
        // what gcc on x86-64 does with the values C leaves undefined: narrow targets go through int
        // and a negative value to an unsigned target through the signed conversion (both modular)
        if constexpr (sizeof (To) < 4) return (To) (int) x ;
        else if constexpr (std::is_unsigned<To>::value) return (x < 0) ? (To) (long long) x : (To) x ;
        else return (To) x ;
    }
    else return (To) x ;
}

template <class To, class From>
__global__ void cast_kernel (const From *__restrict__ in, To *__restrict__ out, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < n ;
        t += (int64_t) gridDim.x * blockDim.x) out [t] = cast_one<To, From> (in [t]) ;
}

template <class To>
static void cast_to (const void *in, int from, void *out, int64_t n, cudaStream_t st)
{
    const int g = grid_for (n) ;
    switch (from)
    {
        case GB200_BOOL   : cast_kernel<To, bool>     <<<g, 256, 0, st>>> ((const bool *) in, (To *) out, n) ; break ;
        case GB200_INT8   : cast_kernel<To, int8_t>   <<<g, 256, 0, st>>> ((const int8_t *) in, (To *) out, n) ; break ;
        case GB200_UINT8  : cast_kernel<To, uint8_t>  <<<g, 256, 0, st>>> ((const uint8_t *) in, (To *) out, n) ; break ;
        case GB200_INT16  : cast_kernel<To, int16_t>  <<<g, 256, 0, st>>> ((const int16_t *) in, (To *) out, n) ; break ;
        case GB200_UINT16 : cast_kernel<To, uint16_t> <<<g, 256, 0, st>>> ((const uint16_t *) in, (To *) out, n) ; break ;
        case GB200_INT32  : cast_kernel<To, int32_t>  <<<g, 256, 0, st>>> ((const int32_t *) in, (To *) out, n) ; break ;
        case GB200_UINT32 : cast_kernel<To, uint32_t> <<<g, 256, 0, st>>> ((const uint32_t *) in, (To *) out, n) ; break ;
        case GB200_INT64  : cast_kernel<To, int64_t>  <<<g, 256, 0, st>>> ((const int64_t *) in, (To *) out, n) ; break ;
        case GB200_UINT64 : cast_kernel<To, uint64_t> <<<g, 256, 0, st>>> ((const uint64_t *) in, (To *) out, n) ; break ;
        case GB200_FP32   : cast_kernel<To, float>    <<<g, 256, 0, st>>> ((const float *) in, (To *) out, n) ; break ;
        default           : cast_kernel<To, double>   <<<g, 256, 0, st>>> ((const double *) in, (To *) out, n) ; break ;
    }
}

// out = the n values `in` of type `from`, cast to type `to` (a new buffer)
gb200_status cast_values (const void *in, int from, int to, int64_t n, DevBuf &out)
{
    Ctx &c = ctx () ;
    GB200_TRY (out.alloc ((size_t) (n > 0 ? n : 1) * type_size (to))) ;
    if (n <= 0) return GB200_SUCCESS ;
    switch (to)
    {
        case GB200_BOOL   : cast_to<bool>     (in, from, out.ptr, n, c.stream) ; break ;
        case GB200_INT8   : cast_to<int8_t>   (in, from, out.ptr, n, c.stream) ; break ;
        case GB200_UINT8  : cast_to<uint8_t>  (in, from, out.ptr, n, c.stream) ; break ;
        case GB200_INT16  : cast_to<int16_t>  (in, from, out.ptr, n, c.stream) ; break ;
        case GB200_UINT16 : cast_to<uint16_t> (in, from, out.ptr, n, c.stream) ; break ;
        case GB200_INT32  : cast_to<int32_t>  (in, from, out.ptr, n, c.stream) ; break ;
        case GB200_UINT32 : cast_to<uint32_t> (in, from, out.ptr, n, c.stream) ; break ;
        case GB200_INT64  : cast_to<int64_t>  (in, from, out.ptr, n, c.stream) ; break ;
        case GB200_UINT64 : cast_to<uint64_t> (in, from, out.ptr, n, c.stream) ; break ;
        case GB200_FP32   : cast_to<float>    (in, from, out.ptr, n, c.stream) ; break ;
        default           : cast_to<double>   (in, from, out.ptr, n, c.stream) ; break ;
    }
    count_launch () ;
    GB200_CUDA (cudaGetLastError ()) ;
    return GB200_SUCCESS ;
}

// ---------------------------------------------------------------------------------------------
// valued masks: an entry of M whose value casts to false does not admit (i,j)
// (reference Gustavson_mask.c:190-197, dot_mask.c:139-141, dot_compmask.c:101-107).  The kernels
// only understand structural masks, so a mask with false-valued entries is filtered once here.
// ---------------------------------------------------------------------------------------------
__global__ void mask_keep_kernel (const void *__restrict__ x, int type_code, int64_t n,
    uint8_t *__restrict__ keep, unsigned long long *__restrict__ nfalse)
{
    unsigned long long local = 0 ;
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < n ;
        t += (int64_t) gridDim.x * blockDim.x)
    {
        bool v ;
        switch (type_code)
        {
            case GB200_BOOL: case GB200_INT8: case GB200_UINT8: v = ((const uint8_t *) x) [t] != 0 ; break ;
            case GB200_INT16: case GB200_UINT16: v = ((const uint16_t *) x) [t] != 0 ; break ;
            case GB200_INT32: case GB200_UINT32: v = ((const uint32_t *) x) [t] != 0 ; break ;
            case GB200_INT64: case GB200_UINT64: v = ((const uint64_t *) x) [t] != 0 ; break ;
            case GB200_FP32: v = ((const float *) x) [t] != 0 ; break ;       // NaN -> true, -0 -> false
            default: v = ((const double *) x) [t] != 0 ; break ;
        }
        keep [t] = v ? 1 : 0 ;
        local += v ? 0 : 1 ;
    }
    if (local) atomicAdd (nfalse, local) ;
}

__global__ void mask_compact_kernel (const int32_t *__restrict__ Mi, const uint8_t *__restrict__ keep,
    const int64_t *__restrict__ pos, int64_t n, int32_t *__restrict__ Mi2)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < n ;
        t += (int64_t) gridDim.x * blockDim.x)
        if (keep [t]) Mi2 [pos [t]] = Mi [t] ;
}

__global__ void remap_ptr_kernel (const int64_t *__restrict__ p, const int64_t *__restrict__ pos,
    int64_t nvec, int64_t *__restrict__ p2)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t <= nvec ;
        t += (int64_t) gridDim.x * blockDim.x) p2 [t] = pos [p [t]] ;
}

// are all stored values equal (bytewise) to the first one
template <class W>
__global__ void iso_kernel (const W *__restrict__ x, int64_t n, unsigned int *__restrict__ differs)
{
    const W x0 = x [0] ;
    bool d = false ;
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < n ;
        t += (int64_t) gridDim.x * blockDim.x) d |= (x [t] != x0) ;
    if (d) *differs = 1u ;
}

gb200_status ensure_iso (gb200_dmatrix_s *d)
{
    if (d->iso_known) return GB200_SUCCESS ;
    Ctx &c = ctx () ;
    const int64_t n = d->v.nnz ;
    d->v.iso = 0 ;
    if (n > 0)
    {
        DevBuf flag ;
        GB200_TRY (flag.alloc (8)) ;
        GB200_CUDA (cudaMemsetAsync (flag.ptr, 0, 8, c.stream)) ;
        const int g = grid_for (n) ;
        switch (type_size (d->v.type_code))
        {
            case 1 : iso_kernel<uint8_t>  <<<g, 256, 0, c.stream>>> ((const uint8_t *)  d->v.x, n, flag.as<unsigned int> ()) ; break ;
            case 2 : iso_kernel<uint16_t> <<<g, 256, 0, c.stream>>> ((const uint16_t *) d->v.x, n, flag.as<unsigned int> ()) ; break ;
            case 4 : iso_kernel<uint32_t> <<<g, 256, 0, c.stream>>> ((const uint32_t *) d->v.x, n, flag.as<unsigned int> ()) ; break ;
            default: iso_kernel<uint64_t> <<<g, 256, 0, c.stream>>> ((const uint64_t *) d->v.x, n, flag.as<unsigned int> ()) ; break ;
        }
        count_launch () ;
        GB200_CUDA (cudaGetLastError ()) ;
        int64_t differs = 0 ;
        GB200_TRY (read_i64 (flag.as<int64_t> (), &differs)) ;
        d->v.iso = (differs == 0) ? 1 : 0 ;
    }
    d->iso_known = 1 ;
    return GB200_SUCCESS ;
}

gb200_status filter_mask (const gb200_dmatrix_s *M, DMat &Mview, DevBuf &Mp2, DevBuf &Mi2)
{
    Ctx &c = ctx () ;
    Mview = M->v ;
    const int64_t n = M->v.nnz ;
    if (n == 0) return GB200_SUCCESS ;
    DevBuf keep, cnt, pos ;
    GB200_TRY (keep.alloc (n)) ;
    GB200_TRY (cnt.alloc (8)) ;
    GB200_CUDA (cudaMemsetAsync (cnt.ptr, 0, 8, c.stream)) ;
    mask_keep_kernel <<<grid_for (n), 256, 0, c.stream>>> (M->v.x, M->v.type_code, n,
        keep.as<uint8_t> (), cnt.as<unsigned long long> ()) ;
    count_launch () ;
    int64_t nfalse = 0 ;
    GB200_TRY (read_i64 (cnt.as<int64_t> (), &nfalse)) ;
    if (nfalse == 0) return GB200_SUCCESS ;
    GB200_TRY (pos.alloc ((n + 1) * sizeof (int64_t))) ;
    GB200_TRY (scan_u8 (keep.as<uint8_t> (), pos.as<int64_t> (), n)) ;
    GB200_TRY (Mi2.alloc ((n - nfalse) * sizeof (int32_t))) ;
    GB200_TRY (Mp2.alloc ((M->v.nvec + 1) * sizeof (int64_t))) ;
    mask_compact_kernel <<<grid_for (n), 256, 0, c.stream>>> (M->v.i, keep.as<uint8_t> (),
        pos.as<int64_t> (), n, Mi2.as<int32_t> ()) ;
    remap_ptr_kernel <<<grid_for (M->v.nvec + 1), 256, 0, c.stream>>> (M->v.p, pos.as<int64_t> (),
        M->v.nvec, Mp2.as<int64_t> ()) ;
    count_launch (2) ;
    GB200_CUDA (cudaGetLastError ()) ;
    Mview.p = Mp2.as<int64_t> () ;
    Mview.i = Mi2.as<int32_t> () ;
    Mview.nnz = n - nfalse ;
    Mview.x = nullptr ;
    Mview.iso = 0 ;
    return GB200_SUCCESS ;
}

// ---------------------------------------------------------------------------------------------
// accumulator -> value conversion
// ---------------------------------------------------------------------------------------------
__global__ void convert_acc_kernel (const uint32_t *__restrict__ acc, void *__restrict__ z,
    int zsize, int is_bool, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < n ;
        t += (int64_t) gridDim.x * blockDim.x)
    {
        const uint32_t a = acc [t] ;
        if (is_bool) ((uint8_t *) z) [t] = (a != 0) ? 1 : 0 ;
        else if (zsize == 1) ((uint8_t *) z) [t] = (uint8_t) a ;
        else ((uint16_t *) z) [t] = (uint16_t) a ;
    }
}

gb200_status convert_acc (const void *acc, int acc_size, void *z, int z_code, int64_t n)
{
    Ctx &c = ctx () ;
    if (n <= 0) return GB200_SUCCESS ;
    const int zsize = type_size (z_code) ;
    if (zsize >= 4)
    {
        GB200_CUDA (cudaMemcpyAsync (z, acc, (size_t) n * zsize, cudaMemcpyDeviceToDevice, c.stream)) ;
        return GB200_SUCCESS ;
    }
    convert_acc_kernel <<<grid_for (n), 256, 0, c.stream>>> ((const uint32_t *) acc, z, zsize,
        z_code == GB200_BOOL, n) ;
    count_launch () ;
    GB200_CUDA (cudaGetLastError ()) ;
    return GB200_SUCCESS ;
}

// ---------------------------------------------------------------------------------------------
// result assembly: hypersparse rule of reference Source/GB_AxB_alloc.c:49-50 and the vector
// bookkeeping of GB_jstartup/jappend/jwrapup (Source/GB.h:4274-4437)
// ---------------------------------------------------------------------------------------------
__global__ void nonempty_kernel (const int64_t *__restrict__ cum, int64_t nsrc, uint8_t *__restrict__ keep)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < nsrc ;
        t += (int64_t) gridDim.x * blockDim.x) keep [t] = (cum [t+1] > cum [t]) ? 1 : 0 ;
}

__global__ void hyper_pack_kernel (const int64_t *__restrict__ cum, const int64_t *__restrict__ names,
    const uint8_t *__restrict__ keep, const int64_t *__restrict__ pos, int64_t nsrc, int64_t cnz,
    int64_t *__restrict__ Ch, int64_t *__restrict__ Cp)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t <= nsrc ;
        t += (int64_t) gridDim.x * blockDim.x)
    {
        if (t == nsrc) { Cp [pos [nsrc]] = cnz ; continue ; }
        if (keep [t])
        {
            Ch [pos [t]] = names ? names [t] : t ;
            Cp [pos [t]] = cum [t] ;
        }
    }
}

// standard-form C from hypersparse sources: Cp[j+1] = entries up to and including vector j
__global__ void scatter_counts_kernel (const int64_t *__restrict__ cum, const int64_t *__restrict__ names,
    int64_t nsrc, int64_t *__restrict__ cntfull)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < nsrc ;
        t += (int64_t) gridDim.x * blockDim.x) cntfull [names [t]] = cum [t+1] - cum [t] ;
}

gb200_status assemble (gb200_result_s *R, int64_t nsrc, const int64_t *names, bool src_hyper,
    DevBuf &cum, DevBuf &Ci, DevBuf &Cx, int64_t cnz, bool C_is_hyper, int64_t cvlen, int64_t cvdim)
{
    Ctx &c = ctx () ;
    R->info.vlen = cvlen ; R->info.vdim = cvdim ; R->info.nnz = cnz ;
    R->info.is_hyper = C_is_hyper ? 1 : 0 ;
    // count the non-empty vectors (always needed: nvec_nonempty must be exact)
    DevBuf keep, pos ;
    GB200_TRY (keep.alloc (nsrc > 0 ? nsrc : 1)) ;
    GB200_TRY (pos.alloc ((nsrc + 1) * sizeof (int64_t))) ;
    if (nsrc > 0)
    {
        nonempty_kernel <<<grid_for (nsrc), 256, 0, c.stream>>> (cum.as<int64_t> (), nsrc, keep.as<uint8_t> ()) ;
        count_launch () ;
    }
    GB200_TRY (scan_u8 (keep.as<uint8_t> (), pos.as<int64_t> (), nsrc)) ;
    int64_t nkeep = 0 ;
    GB200_TRY (read_i64 (pos.as<int64_t> () + nsrc, &nkeep)) ;
    R->info.nvec_nonempty = nkeep ;
    if (C_is_hyper)
    {
        GB200_TRY (R->p.alloc ((nkeep + 1) * sizeof (int64_t))) ;
        GB200_TRY (R->h.alloc ((nkeep > 0 ? nkeep : 1) * sizeof (int64_t))) ;
        hyper_pack_kernel <<<grid_for (nsrc + 1), 256, 0, c.stream>>> (cum.as<int64_t> (), names,
            keep.as<uint8_t> (), pos.as<int64_t> (), nsrc, cnz, R->h.as<int64_t> (), R->p.as<int64_t> ()) ;
        count_launch () ;
        R->info.nvec = nkeep ;
    }
    else if (!src_hyper && nsrc == cvdim)
    {
        R->p = std::move (cum) ;
        R->info.nvec = cvdim ;
    }
    else
    {
        // sources are a hypersparse subset of the vectors of a standard-form C
        DevBuf cntfull ;
        GB200_TRY (cntfull.alloc ((cvdim > 0 ? cvdim : 1) * sizeof (int64_t))) ;
        GB200_CUDA (cudaMemsetAsync (cntfull.ptr, 0, cntfull.bytes, c.stream)) ;
        if (nsrc > 0)
        {
            scatter_counts_kernel <<<grid_for (nsrc), 256, 0, c.stream>>> (cum.as<int64_t> (), names,
                nsrc, cntfull.as<int64_t> ()) ;
            count_launch () ;
        }
        GB200_TRY (R->p.alloc ((cvdim + 1) * sizeof (int64_t))) ;
        GB200_TRY (scan_i64 (cntfull.as<int64_t> (), R->p.as<int64_t> (), cvdim)) ;
        R->info.nvec = cvdim ;
        GB200_CUDA (cudaStreamSynchronize (c.stream)) ;     // cntfull dies at scope exit
    }
    GB200_CUDA (cudaGetLastError ()) ;
    R->i = std::move (Ci) ;
    R->x = std::move (Cx) ;
    return GB200_SUCCESS ;
}

} // namespace gb200

// =============================================================================================
// C ABI (part 1)
// =============================================================================================
using namespace gb200 ;

extern "C" {
#pragma GCC visibility push(default)

const char *gb200_last_error (void) { return g_err ; }
const char *gb200_version (void) { return "gb_b200 0.1 (sm_100a; reference: SuiteSparse:GraphBLAS 2.3.3)" ; }

int gb200_device_count (void)
{
    int n = 0 ;
    if (cudaGetDeviceCount (&n) != cudaSuccess) { cudaGetLastError () ; return 0 ; }
    return n ;
}

gb200_status gb200_init (int device) { return do_init (device) ; }

gb200_status gb200_finalize (void)
{
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    if (!c.ready) return GB200_SUCCESS ;
    cudaStreamSynchronize (c.stream) ;
    dev_pool_trim () ;
    cudaFreeHost (c.pinned) ; c.pinned = nullptr ;
    cudaEventDestroy (c.ev0) ; cudaEventDestroy (c.ev1) ;
    cudaEventDestroy (c.fork_ev) ; c.fork_ev = nullptr ;
    for (int k = 0 ; k < Ctx::NSIDE ; k++)
    {
        cudaStreamSynchronize (c.side [k]) ;
        cudaStreamDestroy (c.side [k]) ; c.side [k] = nullptr ;
        cudaEventDestroy (c.side_done [k]) ; c.side_done [k] = nullptr ;
    }
    for (cudaEvent_t e : c.kev) cudaEventDestroy (e) ;
    c.kev.clear () ; c.kev_used = 0 ;
    cudaStreamDestroy (c.stream) ; c.stream = nullptr ;
    c.ready = false ;
    return GB200_SUCCESS ;
}

void gb200_device_trim (void)
{
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    if (c.ready) dev_pool_trim () ;
}

int64_t gb200_kernel_launches (void) { return ctx ().launches.load () ; }
int64_t gb200_multiplies (void) { return ctx ().multiplies.load () ; }

// ---- semiring canonicalisation: reference Source/GB_semiring_builtin.c:86-148 and
//      Source/GB_boolean_rename.c:30-91 -------------------------------------------------------
static int boolean_rename (int op)
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

gb200_status gb200_semiring_canonical (gb200_semiring *s)
{
    if (s == NULL) return GB200_INVALID ;
    if (s->xy_code < GB200_BOOL || s->xy_code > GB200_FP64 || s->z_code < GB200_BOOL
        || s->z_code > GB200_FP64 || s->mult_opcode < GB200_FIRST || s->mult_opcode > GB200_LE
        || s->add_opcode < GB200_FIRST || s->add_opcode > GB200_LE)
    {
        set_error ("semiring outside the built-in operator/type space (user-defined?)") ;
        return GB200_NOT_SUPPORTED ;
    }
    if (s->xy_code == GB200_BOOL) s->mult_opcode = boolean_rename (s->mult_opcode) ;
    if (s->z_code == GB200_BOOL) s->add_opcode = boolean_rename (s->add_opcode) ;
    if (s->flipxy)
    {
        switch (s->mult_opcode)
        {
            case GB200_FIRST  : s->mult_opcode = GB200_SECOND ; break ;
            case GB200_SECOND : s->mult_opcode = GB200_FIRST ; break ;
            case GB200_GT : s->mult_opcode = GB200_LT ; break ;
            case GB200_LT : s->mult_opcode = GB200_GT ; break ;
            case GB200_GE : s->mult_opcode = GB200_LE ; break ;
            case GB200_LE : s->mult_opcode = GB200_GE ; break ;
            case GB200_ISGT : s->mult_opcode = GB200_ISLT ; break ;
            case GB200_ISLT : s->mult_opcode = GB200_ISGT ; break ;
            case GB200_ISGE : s->mult_opcode = GB200_ISLE ; break ;
            case GB200_ISLE : s->mult_opcode = GB200_ISGE ; break ;
            default : break ;
        }
    }
    const bool mult_is_compare = (s->mult_opcode >= GB200_EQ) ;
    const bool zbool = (s->z_code == GB200_BOOL) ;
    bool ok ;
    if (mult_is_compare) ok = zbool ;                       // TxT -> bool
    else ok = (s->z_code == s->xy_code) ;                    // TxT -> T
    if (zbool)
        ok = ok && (s->add_opcode == GB200_LOR || s->add_opcode == GB200_LAND
            || s->add_opcode == GB200_LXOR || s->add_opcode == GB200_EQ) ;
    else
        ok = ok && (s->add_opcode == GB200_MIN || s->add_opcode == GB200_MAX
            || s->add_opcode == GB200_PLUS || s->add_opcode == GB200_TIMES) ;
    if (s->xy_code == GB200_BOOL)
    {
        const int m = s->mult_opcode ;
        ok = ok && (m == GB200_FIRST || m == GB200_SECOND || m == GB200_LOR || m == GB200_LAND
            || m == GB200_LXOR || m == GB200_EQ || m == GB200_GT || m == GB200_LT || m == GB200_GE
            || m == GB200_LE) ;
    }
    if (!ok)
    {
        set_error ("semiring (add %d, mult %d, xy %d, z %d) is not one of the built-in workers",
            s->add_opcode, s->mult_opcode, s->xy_code, s->z_code) ;
        return GB200_NOT_SUPPORTED ;
    }
    return GB200_SUCCESS ;
}

// ---- upload ------------------------------------------------------------------------------------
static gb200_status upload_any (gb200_dmatrix *out, const gb200_matrix *host, int64_t nnz_known)
{
    if (out == NULL || host == NULL) return GB200_INVALID ;
    *out = NULL ;
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    if (host->vlen < 0 || host->vdim < 0 || host->nvec < 0 || host->p == NULL
        || host->type_code < GB200_BOOL || host->type_code > GB200_FP64
        || (host->h == NULL && host->nvec != host->vdim) || host->nvec > host->vdim)
    {
        set_error ("gb200_upload: malformed matrix") ;
        return GB200_INVALID ;
    }
    if (host->vlen >= (int64_t) INT32_MAX || host->vdim >= (int64_t) INT32_MAX)
    {
        set_error ("gb200_upload: dimension %lld x %lld needs 64-bit device indices (not built)",
            (long long) host->vlen, (long long) host->vdim) ;
        return GB200_NOT_SUPPORTED ;
    }
    const int64_t nvec = host->nvec ;
    // nnz_known >= 0: the arrays are device memory (p cannot be read here); the caller vouches for p[0] == 0
    const int64_t nnz = (nnz_known >= 0) ? nnz_known : (host->p [nvec] - host->p [0]) ;
    if ((nnz_known < 0 && host->p [0] != 0) || nnz < 0 || (nnz > 0 && (host->i == NULL || host->x == NULL)))
    {
        set_error ("gb200_upload: malformed matrix (p[0] != 0 or missing arrays)") ;
        return GB200_INVALID ;
    }
    gb200_dmatrix_s *d = new (std::nothrow) gb200_dmatrix_s () ;
    if (d == NULL) return GB200_OUT_OF_MEMORY ;
    gb200_status st = GB200_SUCCESS ;
    const int tsz = type_size (host->type_code) ;
    auto body = [&] () -> gb200_status
    {
        GB200_TRY (d->p.alloc ((nvec + 1) * sizeof (int64_t))) ;
        GB200_CUDA (cudaMemcpyAsync (d->p.ptr, host->p, (nvec + 1) * sizeof (int64_t),
            cudaMemcpyDefault, c.stream)) ;
        if (host->h != NULL)
        {
            GB200_TRY (d->h.alloc ((nvec > 0 ? nvec : 1) * sizeof (int64_t))) ;
            if (nvec > 0)
                GB200_CUDA (cudaMemcpyAsync (d->h.ptr, host->h, nvec * sizeof (int64_t),
                    cudaMemcpyDefault, c.stream)) ;
        }
        // 32 bytes of slack: the masked dot kernel reads indices in aligned 16- or 32-byte chunks
        GB200_TRY (d->i.alloc (((nnz > 0 ? nnz : 1) + 8) * sizeof (int32_t))) ;
        GB200_CUDA (cudaMemsetAsync (d->i.as<int32_t> () + (nnz > 0 ? nnz : 1), 0, 8 * sizeof (int32_t), c.stream)) ;
        GB200_TRY (d->x.alloc ((nnz > 0 ? nnz : 1) * (size_t) tsz)) ;
        if (nnz > 0)
        {
            // 64-bit indices are staged in chunks and narrowed on the device
            const int64_t chunk = (nnz < (1LL << 26)) ? nnz : (1LL << 26) ;
            DevBuf stage ;
            GB200_TRY (stage.alloc (chunk * sizeof (int64_t))) ;
            for (int64_t off = 0 ; off < nnz ; off += chunk)
            {
                const int64_t len = (nnz - off < chunk) ? (nnz - off) : chunk ;
                GB200_CUDA (cudaMemcpyAsync (stage.ptr, host->i + off, len * sizeof (int64_t),
                    cudaMemcpyDefault, c.stream)) ;
                narrow_idx_kernel <<<grid_for (len), 256, 0, c.stream>>> (stage.as<int64_t> (),
                    d->i.as<int32_t> () + off, len) ;
                count_launch () ;
            }
            GB200_CUDA (cudaMemcpyAsync (d->x.ptr, host->x, (size_t) nnz * tsz,
                cudaMemcpyDefault, c.stream)) ;
            GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
        }
        GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
        GB200_CUDA (cudaGetLastError ()) ;
        return GB200_SUCCESS ;
    } ;
    st = body () ;
    if (st != GB200_SUCCESS) { delete d ; return st ; }
    d->is_hyper_flag = (host->h != NULL) ;
    d->v.p = d->p.as<int64_t> () ;
    d->v.h = (host->h != NULL) ? d->h.as<int64_t> () : nullptr ;
    d->v.i = d->i.as<int32_t> () ;
    d->v.x = d->x.ptr ;
    d->v.vlen = host->vlen ; d->v.vdim = host->vdim ; d->v.nvec = nvec ; d->v.nnz = nnz ;
    d->v.hyper = (host->h != NULL && nvec < host->vdim) ? 1 : 0 ;
    d->v.type_code = host->type_code ;
    d->v.iso = 0 ;
    *out = d ;
    return GB200_SUCCESS ;
}

gb200_status gb200_upload (gb200_dmatrix *out, const gb200_matrix *host)
{
    return upload_any (out, host, -1) ;
}

gb200_status gb200_upload_from_device (gb200_dmatrix *out, const gb200_matrix *dev, int64_t nnz)
{
    if (nnz < 0) return GB200_INVALID ;
    return upload_any (out, dev, nnz) ;
}

gb200_status gb200_dmatrix_free (gb200_dmatrix *d)
{
    if (d == NULL || *d == NULL) return GB200_SUCCESS ;
    std::lock_guard<std::recursive_mutex> lock (ctx ().mu) ;
    delete *d ;
    *d = NULL ;
    return GB200_SUCCESS ;
}

// ---- results -----------------------------------------------------------------------------------
gb200_status gb200_result_get_info (gb200_result r, gb200_result_info *info)
{
    if (r == NULL || info == NULL) return GB200_INVALID ;
    *info = r->info ;
    return GB200_SUCCESS ;
}

gb200_status gb200_result_fetch (gb200_result r, int64_t *p, int64_t *h, int64_t *i, void *x)
{
    if (r == NULL || p == NULL) return GB200_INVALID ;
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    const gb200_result_info &f = r->info ;
    // every argument is checked before any copy is queued
    if (f.nnz > 0 && (i == NULL || x == NULL)) { set_error ("gb200_result_fetch: NULL i or x") ; return GB200_INVALID ; }
    GB200_CUDA (cudaMemcpyAsync (p, r->p.ptr, (f.nvec + 1) * sizeof (int64_t),
        cudaMemcpyDefault, c.stream)) ;
    if (f.is_hyper && h != NULL && f.nvec > 0)
        GB200_CUDA (cudaMemcpyAsync (h, r->h.ptr, f.nvec * sizeof (int64_t),
            cudaMemcpyDefault, c.stream)) ;
    if (f.nnz > 0)
    {
        const int64_t chunk = (f.nnz < (1LL << 26)) ? f.nnz : (1LL << 26) ;
        DevBuf stage ;
        GB200_TRY (stage.alloc (chunk * sizeof (int64_t))) ;
        for (int64_t off = 0 ; off < f.nnz ; off += chunk)
        {
            const int64_t len = (f.nnz - off < chunk) ? (f.nnz - off) : chunk ;
            widen_idx_kernel <<<grid_for (len), 256, 0, c.stream>>> (r->i.as<int32_t> () + off,
                stage.as<int64_t> (), len) ;
            count_launch () ;
            GB200_CUDA (cudaMemcpyAsync (i + off, stage.ptr, len * sizeof (int64_t),
                cudaMemcpyDefault, c.stream)) ;
        }
        GB200_CUDA (cudaMemcpyAsync (x, r->x.ptr, (size_t) f.nnz * type_size (f.type_code),
            cudaMemcpyDefault, c.stream)) ;
        GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
    }
    GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
    return GB200_SUCCESS ;
}

gb200_status gb200_result_free (gb200_result *r)
{
    if (r == NULL || *r == NULL) return GB200_SUCCESS ;
    std::lock_guard<std::recursive_mutex> lock (ctx ().mu) ;
    delete *r ;
    *r = NULL ;
    return GB200_SUCCESS ;
}

gb200_status gb200_partition_by_flops (const int64_t *cum, int64_t nvec, int nparts, int64_t *bounds)
{
    if (cum == NULL || bounds == NULL || nparts < 1 || nvec < 0) return GB200_INVALID ;
    const int64_t total = cum [nvec] ;
    bounds [0] = 0 ;
    for (int g = 1 ; g < nparts ; g++)
    {
        // first vector boundary whose cumulative flops reach g/nparts of the total
        const double target = (double) total * (double) g / (double) nparts ;
        int64_t lo = bounds [g-1], hi = nvec ;
        while (lo < hi)
        {
            int64_t mid = (lo + hi) >> 1 ;
            if ((double) cum [mid] < target) lo = mid + 1 ; else hi = mid ;
        }
        bounds [g] = lo ;
    }
    bounds [nparts] = nvec ;
    return GB200_SUCCESS ;
}

#pragma GCC visibility pop
} // extern "C"
