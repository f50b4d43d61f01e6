// engine_host.cu -- host-side services of the C ABI that are not part of a multiply:
//   * the page-locked malloc/calloc/realloc/free quartet a host application hands to GxB_init
//     (reference Include/GraphBLAS.h:330-340; every GraphBLAS array then lives in pinned memory,
//     including the T that the shim allocates through GB_create, Source/GB.h:1021-1035)
//   * CUDA-event timers on the library's stream
#include <cstdlib>
#include <cstring>
#include <map>
#include <unordered_map>
#include <vector>
#include <unistd.h>
#include "engine.cuh"

namespace gb200 {

// ---------------------------------------------------------------------------------------------
// pinned block cache.  Page-locking is expensive, so freed blocks are kept by capacity (rounded up
// to a size class: powers of two split in four) and handed out again; a block is reused only for a
// request that fills at least half of it.
// ---------------------------------------------------------------------------------------------
struct HostPool
{
    std::mutex mu ;
    std::unordered_map<void *, size_t> live ;           // pinned blocks in use: ptr -> capacity
    std::multimap<size_t, void *> cache ;               // free pinned blocks by capacity
    size_t cached_bytes = 0 ;
    size_t cache_limit = 0 ;                            // most bytes kept in `cache` (0: not read yet)
    bool no_device = false ;                            // there is no CUDA device at all: plain malloc
} ;

static HostPool &pool () { static HostPool *p = new HostPool () ; return *p ; }    // never destroyed

// Upper bound of the free-block cache: GB200_HOST_CACHE_MB, else a quarter of the machine's RAM (a
// long-running host with varying sizes must not accumulate page-locked memory without limit).
static size_t cache_limit_locked (HostPool &hp)
{
    if (hp.cache_limit == 0)
    {
        const char *env = getenv ("GB200_HOST_CACHE_MB") ;
        if (env != nullptr && atoll (env) >= 0) hp.cache_limit = ((size_t) atoll (env) << 20) + 1 ;
        else
        {
            const long pages = sysconf (_SC_PHYS_PAGES), psz = sysconf (_SC_PAGE_SIZE) ;
            size_t ram = (pages > 0 && psz > 0) ? (size_t) pages * (size_t) psz : ((size_t) 16 << 30) ;
            hp.cache_limit = ram / 4 + 1 ;
        }
    }
    return hp.cache_limit ;
}

// evict largest-first until the cache fits its limit; the evicted blocks are unpinned by the caller
// outside the lock (cudaFreeHost synchronises)
static void evict_locked (HostPool &hp, std::vector<void *> &drop)
{
    const size_t limit = cache_limit_locked (hp) ;
    while (hp.cached_bytes > limit && !hp.cache.empty ())
    {
        auto it = std::prev (hp.cache.end ()) ;
        hp.cached_bytes -= it->first ;
        drop.push_back (it->second) ;
        hp.cache.erase (it) ;
    }
}

static size_t size_class (size_t n)
{
    size_t c = GB200_HOST_PIN_MIN ;
    while (c < n) c <<= 1 ;
    if (c == n || c == GB200_HOST_PIN_MIN) return c ;
    // four classes per octave: 1.25, 1.5, 1.75, 2.0 x the lower power of two
    const size_t lo = c >> 1, step = lo >> 2 ;
    for (size_t q = lo + step ; q < c ; q += step) if (q >= n) return q ;
    return c ;
}

static void *pinned_get (size_t size)
{
    HostPool &hp = pool () ;
    const size_t cap = size_class (size) ;
    {
        std::lock_guard<std::mutex> lock (hp.mu) ;
        if (hp.no_device) return nullptr ;
        auto it = hp.cache.lower_bound (cap) ;
        if (it != hp.cache.end () && it->first <= 2 * cap)
        {
            void *p = it->second ;
            const size_t c = it->first ;
            hp.cache.erase (it) ;
            hp.cached_bytes -= c ;
            hp.live [p] = c ;
            return p ;
        }
    }
    void *p = nullptr ;
    cudaError_t e = cudaMallocHost (&p, cap) ;
    if (e != cudaSuccess)
    {
        cudaGetLastError () ;
        // out of pinnable memory: drop the cache and retry once.  No device / no driver: plain malloc
        // from now on.  Any other (possibly transient) error: plain malloc for this block only.
        if (e == cudaErrorMemoryAllocation)
        {
            gb200_host_trim () ;
            e = cudaMallocHost (&p, cap) ;
            if (e != cudaSuccess) { cudaGetLastError () ; return nullptr ; }
        }
        else
        {
            if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
            {
                std::lock_guard<std::mutex> lock (hp.mu) ;
                hp.no_device = true ;
            }
            return nullptr ;
        }
    }
    std::lock_guard<std::mutex> lock (hp.mu) ;
    hp.live [p] = cap ;
    return p ;
}

} // namespace gb200

using namespace gb200 ;

extern "C" {
#pragma GCC visibility push(default)

void *gb200_host_malloc (size_t size)
{
    if (size >= GB200_HOST_PIN_MIN)
    {
        void *p = pinned_get (size) ;
        if (p != nullptr) return p ;
    }
    return malloc (size > 0 ? size : 1) ;
}

void *gb200_host_calloc (size_t n, size_t size)
{
    if (size != 0 && n > SIZE_MAX / size) return nullptr ;
    const size_t bytes = n * size ;
    if (bytes >= GB200_HOST_PIN_MIN)
    {
        void *p = pinned_get (bytes) ;
        if (p != nullptr) { memset (p, 0, bytes) ; return p ; }
    }
    return calloc (n > 0 ? n : 1, size > 0 ? size : 1) ;
}

void gb200_host_free (void *p)
{
    if (p == nullptr) return ;
    gb200_cache_invalidate (p) ;        // a device copy made from this array is stale now
    HostPool &hp = pool () ;
    std::vector<void *> drop ;
    bool mine = false ;
    {
        std::lock_guard<std::mutex> lock (hp.mu) ;
        auto it = hp.live.find (p) ;
        if (it != hp.live.end ())
        {
            const size_t cap = it->second ;
            hp.live.erase (it) ;
            hp.cache.emplace (cap, p) ;
            hp.cached_bytes += cap ;
            evict_locked (hp, drop) ;
            mine = true ;
        }
    }
    for (void *q : drop) cudaFreeHost (q) ;
    if (!drop.empty ()) cudaGetLastError () ;
    if (!mine) free (p) ;
}

void *gb200_host_realloc (void *p, size_t size)
{
    if (p == nullptr) return gb200_host_malloc (size) ;
    gb200_cache_invalidate (p) ;        // the array moves or changes length
    HostPool &hp = pool () ;
    size_t cap = 0 ;
    {
        std::lock_guard<std::mutex> lock (hp.mu) ;
        auto it = hp.live.find (p) ;
        if (it != hp.live.end ()) cap = it->second ;
    }
    if (cap == 0)
    {
        // a malloc block: stays one unless it grows past the pinning threshold
        if (size < GB200_HOST_PIN_MIN) return realloc (p, size > 0 ? size : 1) ;
        // its old size is unknown to us, so let realloc move it first, then copy into a pinned block
        void *q = realloc (p, size) ;
        if (q == nullptr) return nullptr ;
        void *r = pinned_get (size) ;
        if (r == nullptr) return q ;
        memcpy (r, q, size) ;
        free (q) ;
        return r ;
    }
    if (size <= cap && size >= cap / 4) return p ;         // fits: keep the block
    void *q = gb200_host_malloc (size) ;
    if (q == nullptr) return nullptr ;
    memcpy (q, p, size < cap ? size : cap) ;
    gb200_host_free (p) ;
    return q ;
}

void gb200_host_trim (void)
{
    HostPool &hp = pool () ;
    std::multimap<size_t, void *> drop ;
    {
        std::lock_guard<std::mutex> lock (hp.mu) ;
        drop.swap (hp.cache) ;
        hp.cached_bytes = 0 ;
    }
    for (auto &kv : drop) cudaFreeHost (kv.second) ;
    cudaGetLastError () ;
}

// ---- timers ------------------------------------------------------------------------------------
static cudaEvent_t g_timer [8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr } ;

gb200_status gb200_timer_mark (int slot)
{
    if (slot < 0 || slot >= 8) return GB200_INVALID ;
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    if (g_timer [slot] == nullptr) GB200_CUDA (cudaEventCreate (&g_timer [slot])) ;
    GB200_CUDA (cudaEventRecord (g_timer [slot], c.stream)) ;
    return GB200_SUCCESS ;
}

gb200_status gb200_timer_elapsed_ms (int slot_a, int slot_b, double *ms)
{
    if (slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8 || ms == nullptr) return GB200_INVALID ;
    if (g_timer [slot_a] == nullptr || g_timer [slot_b] == nullptr)
    {
        set_error ("gb200_timer_elapsed_ms: slot not marked") ;
        return GB200_INVALID ;
    }
    GB200_TRY (ensure_init ()) ;
    GB200_CUDA (cudaEventSynchronize (g_timer [slot_b])) ;
    float t = 0 ;
    GB200_CUDA (cudaEventElapsedTime (&t, g_timer [slot_a], g_timer [slot_b])) ;
    *ms = t ;
    return GB200_SUCCESS ;
}

#pragma GCC visibility pop
} // extern "C"
