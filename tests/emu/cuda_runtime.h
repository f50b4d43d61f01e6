// tests/emu/cuda_runtime.h -- a HOST stand-in for the CUDA device environment, for tests only.
//
// tests/test_kernel_emulation.py compiles the kernel SOURCES of graphblas_b200/csrc (kernels.cuh and the
// headers it includes) with g++ and this directory first on the include path, so that
// `#include <cuda_runtime.h>` lands here.  One CUDA thread is one OS thread:
//   * a block is launched as blockDim.x std::threads; blocks of a grid run one after another;
//   * threadIdx is thread_local, blockIdx / blockDim / gridDim are set per block;
//   * __syncthreads is a barrier over the block; warp intrinsics (__shfl*_sync, __ballot_sync,
//     __any_sync, __reduce_add_sync) are a rendezvous of the lanes named by the mask (the full warp, or
//     an aligned lane group) plus an exchange buffer;
//   * __shared__ variables are function-local statics (one block at a time, so one copy is right);
//     the one `extern __shared__` array is rewritten by the test driver to point at emu::dyn_smem;
//   * atomics are the GCC __atomic builtins; __ldg / __ldcs are plain loads.
// Nothing here is shipped or linked into the product.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <thread>
#include <vector>
#include <barrier>
#include <atomic>
#include <memory>
#include <functional>
#include <mutex>
#include <condition_variable>
#include <ctime>
using std::isnan ;
using std::isinf ;

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__ ((always_inline))
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas (n)
#define GB200_HOST_EMULATION 1

struct uint3 { unsigned x, y, z ; } ;
struct dim3 { unsigned x, y, z ; dim3 (unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x (x_), y (y_), z (z_) { } } ;
struct uint2 { unsigned x, y ; } ;
static inline uint2 make_uint2 (unsigned x, unsigned y) { uint2 r = { x, y } ; return r ; }
struct int4 { int x, y, z, w ; } ;
static inline int4 make_int4 (int x, int y, int z, int w) { int4 r = { x, y, z, w } ; return r ; }
typedef void *cudaStream_t ;

namespace emu {
// A rendezvous of the lanes named by `mask` (the full warp, or an aligned group of lanes as dotv_kernel and
// dot_kernel use): a counting barrier per (lowest lane, size) of the mask -- lanes that left a loop early may
// wait on the full mask while a lane group of the same warp still meets on its own mask; waiting lanes yield.
struct Group { std::atomic<int> count { 0 } ; std::atomic<int> gen { 0 } ; } ;
struct Warp
{
    Group grp [32][33] ;
    uint64_t slot [32] ;
} ;
struct Block
{
    std::unique_ptr<std::barrier<>> bar ;
    std::vector<std::unique_ptr<Warp>> warps ;
} ;
inline Block *g_block = nullptr ;
inline thread_local Warp *t_warp = nullptr ;
inline thread_local int t_lane = 0 ;
alignas (128) inline unsigned char dyn_smem [232448] ;          // 227 KB

static inline void group_sync (unsigned mask)
{
    if (t_warp == nullptr) { fprintf (stderr, "emu: warp intrinsic in a kernel launched thread-by-thread\n") ; abort () ; }
    const int leader = __builtin_ffs ((int) mask) - 1, n = __builtin_popcount (mask) ;
    if (!((mask >> t_lane) & 1u)) { fprintf (stderr, "emu: lane %d is not in the mask %08x it syncs on\n", t_lane, mask) ; abort () ; }
    Group &g = t_warp->grp [leader][n] ;
    const int gen = g.gen.load (std::memory_order_acquire) ;
    if (g.count.fetch_add (1, std::memory_order_acq_rel) + 1 == n)
    {
        g.count.store (0, std::memory_order_relaxed) ;
        g.gen.fetch_add (1, std::memory_order_release) ;
    }
    else while (g.gen.load (std::memory_order_acquire) == gen) std::this_thread::yield () ;
}

template <class T> static inline uint64_t bits (T v) { uint64_t b = 0 ; memcpy (&b, &v, sizeof (T)) ; return b ; }
template <class T> static inline T unbits (uint64_t b) { T v ; memcpy (&v, &b, sizeof (T)) ; return v ; }
// every lane of the mask publishes a word, then reads the word of lane `src`
static inline uint64_t exchange (unsigned mask, uint64_t mine, int src)
{
    t_warp->slot [t_lane] = mine ;
    group_sync (mask) ;
    const uint64_t r = t_warp->slot [src & 31] ;
    group_sync (mask) ;
    return r ;
}
} // namespace emu

inline thread_local uint3 threadIdx = { 0, 0, 0 } ;
inline uint3 blockIdx = { 0, 0, 0 } ;
inline dim3 blockDim, gridDim ;

static inline void __syncthreads ()
{
    if (emu::g_block == nullptr) { fprintf (stderr, "emu: __syncthreads in a kernel launched thread-by-thread\n") ; abort () ; }
    emu::g_block->bar->arrive_and_wait () ;
}
static inline void __syncwarp (unsigned mask = 0xffffffffu) { emu::group_sync (mask) ; }
static inline void __threadfence () { __atomic_thread_fence (__ATOMIC_SEQ_CST) ; }
static inline void __threadfence_system () { __atomic_thread_fence (__ATOMIC_SEQ_CST) ; }
static inline void __nanosleep (unsigned) { }
static inline void __trap () { fprintf (stderr, "emu: __trap()\n") ; abort () ; }
static inline int __popc (unsigned v) { return __builtin_popcount (v) ; }
static inline int __ffs (unsigned v) { return __builtin_ffs ((int) v) ; }
template <class T> static inline T __ldg (const T *p) { return *p ; }
template <class T> static inline T __ldcs (const T *p) { return *p ; }

template <class T> static inline T __shfl_sync (unsigned mask, T v, int src, int width = 32)
{
    const int base = emu::t_lane & ~(width - 1) ;
    return emu::unbits<T> (emu::exchange (mask, emu::bits (v), base + (src & (width - 1)))) ;
}
// a source lane outside the mask (possible only at the edge of a group) gives the caller's own value,
// which is what the kernels rely on: they guard the use with `gl + off < G`
template <class T> static inline T __shfl_down_sync (unsigned mask, T v, unsigned off, int width = 32)
{
    const int pos = emu::t_lane & (width - 1) ;
    int src = (pos + (int) off < width) ? (emu::t_lane + (int) off) : emu::t_lane ;
    if (!((mask >> src) & 1u)) src = emu::t_lane ;
    return emu::unbits<T> (emu::exchange (mask, emu::bits (v), src)) ;
}
template <class T> static inline T __shfl_up_sync (unsigned mask, T v, unsigned off, int width = 32)
{
    const int pos = emu::t_lane & (width - 1) ;
    int src = (pos - (int) off >= 0) ? (emu::t_lane - (int) off) : emu::t_lane ;
    if (!((mask >> src) & 1u)) src = emu::t_lane ;
    return emu::unbits<T> (emu::exchange (mask, emu::bits (v), src)) ;
}
template <class T> static inline T __shfl_xor_sync (unsigned mask, T v, int lanemask, int width = 32)
{
    int src = emu::t_lane ^ lanemask ;
    if ((src & ~(width - 1)) != (emu::t_lane & ~(width - 1)) || !((mask >> src) & 1u)) src = emu::t_lane ;
    return emu::unbits<T> (emu::exchange (mask, emu::bits (v), src)) ;
}
static inline unsigned __ballot_sync (unsigned mask, int pred)
{
    emu::t_warp->slot [emu::t_lane] = pred ? 1 : 0 ;
    emu::group_sync (mask) ;
    unsigned m = 0 ;
    for (int l = 0 ; l < 32 ; l++) if (((mask >> l) & 1u) && emu::t_warp->slot [l]) m |= (1u << l) ;
    emu::group_sync (mask) ;
    return m ;
}
static inline int __any_sync (unsigned mask, int pred) { return __ballot_sync (mask, pred) != 0 ; }
static inline int __all_sync (unsigned mask, int pred) { return __ballot_sync (mask, pred) == mask ; }
static inline unsigned __reduce_add_sync (unsigned mask, unsigned v)
{
    emu::t_warp->slot [emu::t_lane] = v ;
    emu::group_sync (mask) ;
    unsigned s = 0 ;
    for (int l = 0 ; l < 32 ; l++) if ((mask >> l) & 1u) s += (unsigned) emu::t_warp->slot [l] ;
    emu::group_sync (mask) ;
    return s ;
}

static inline unsigned __reduce_or_sync (unsigned mask, unsigned v)
{
    emu::t_warp->slot [emu::t_lane] = v ;
    emu::group_sync (mask) ;
    unsigned s = 0 ;
    for (int l = 0 ; l < 32 ; l++) if ((mask >> l) & 1u) s |= (unsigned) emu::t_warp->slot [l] ;
    emu::group_sync (mask) ;
    return s ;
}
static inline int __clz (unsigned v) { return v ? __builtin_clz (v) : 32 ; }

// ---- atomics ----------------------------------------------------------------------------------
#define EMU_ATOMIC_INT(T) \
static inline T atomicAdd (T *p, T v) { return __atomic_fetch_add (p, v, __ATOMIC_RELAXED) ; } \
static inline T atomicExch (T *p, T v) { return __atomic_exchange_n (p, v, __ATOMIC_RELAXED) ; } \
static inline T atomicOr (T *p, T v) { return __atomic_fetch_or (p, v, __ATOMIC_RELAXED) ; } \
static inline T atomicAnd (T *p, T v) { return __atomic_fetch_and (p, v, __ATOMIC_RELAXED) ; } \
static inline T atomicXor (T *p, T v) { return __atomic_fetch_xor (p, v, __ATOMIC_RELAXED) ; } \
static inline T atomicCAS (T *p, T expect, T v) \
{ __atomic_compare_exchange_n (p, &expect, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED) ; return expect ; } \
static inline T atomicMin (T *p, T v) \
{ T old = __atomic_load_n (p, __ATOMIC_RELAXED) ; \
  while (v < old && !__atomic_compare_exchange_n (p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { } return old ; } \
static inline T atomicMax (T *p, T v) \
{ T old = __atomic_load_n (p, __ATOMIC_RELAXED) ; \
  while (v > old && !__atomic_compare_exchange_n (p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { } return old ; }
EMU_ATOMIC_INT (int)
EMU_ATOMIC_INT (unsigned int)
EMU_ATOMIC_INT (unsigned long long)
EMU_ATOMIC_INT (long long)
#undef EMU_ATOMIC_INT
#define EMU_ATOMIC_FP(T, U) \
static inline T atomicAdd (T *p, T v) \
{ U old = __atomic_load_n ((U *) p, __ATOMIC_RELAXED) ; \
  while (true) { T o ; memcpy (&o, &old, sizeof (T)) ; T n = o + v ; U nb ; memcpy (&nb, &n, sizeof (T)) ; \
    if (__atomic_compare_exchange_n ((U *) p, &old, nb, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return o ; } }
EMU_ATOMIC_FP (float, unsigned int)
EMU_ATOMIC_FP (double, unsigned long long)
#undef EMU_ATOMIC_FP

// ---- launching ----------------------------------------------------------------------------------
namespace emu {
// A pool of OS threads that play the CUDA threads of one block at a time (blocks of a grid run one after
// another: the single-pass scan's look-back and the kernels that pull work from a global counter only ever
// wait for EARLIER blocks).  The pool is never destroyed: the process may exit with its threads parked.
struct Pool
{
    std::mutex mu ;
    std::condition_variable go, done ;
    std::vector<std::thread> th ;
    uint64_t gen = 0 ;
    unsigned want = 0, running = 0 ;
    const std::function<void ()> *body = nullptr ;
    Block *blk = nullptr ;
} ;
static inline Pool &pool () { static Pool *p = new Pool () ; return *p ; }

static inline void worker (unsigned t)
{
    Pool &P = pool () ;
    uint64_t seen = 0 ;
    while (true)
    {
        const std::function<void ()> *body ;
        Block *blk ;
        {
            std::unique_lock<std::mutex> lk (P.mu) ;
            P.go.wait (lk, [&] { return P.gen != seen ; }) ;
            seen = P.gen ;
            if (t >= P.want) continue ;
            body = P.body ; blk = P.blk ;
        }
        threadIdx.x = t ; t_lane = (int) (t & 31) ; t_warp = blk->warps [t >> 5].get () ;
        (*body) () ;
        {
            std::lock_guard<std::mutex> lk (P.mu) ;
            if (--P.running == 0) P.done.notify_one () ;
        }
    }
}

// run `body` (a call of one __global__ function) as a grid of `grid` blocks of `block` threads
static inline void launch (dim3 grid, unsigned block, const std::function<void ()> &body)
{
    if (block == 0 || grid.x == 0 || grid.y == 0) return ;
    if (block % 32 != 0) { fprintf (stderr, "emu: block size %u is not a multiple of 32\n", block) ; abort () ; }
    static std::mutex one_launch ;                  // kernels of one stream run one after another
    std::lock_guard<std::mutex> serial (one_launch) ;
    Pool &P = pool () ;
    while (P.th.size () < block) { const unsigned t = (unsigned) P.th.size () ; P.th.emplace_back (worker, t) ; P.th.back ().detach () ; }
    gridDim = grid ; blockDim.x = block ;
    Block blk ;
    blk.bar.reset (new std::barrier<> (block)) ;
    for (unsigned w = 0 ; w < block / 32 ; w++) blk.warps.emplace_back (new Warp ()) ;
    g_block = &blk ;
    for (unsigned by = 0 ; by < grid.y ; by++)
    for (unsigned b = 0 ; b < grid.x ; b++)
    {
        blockIdx.x = b ; blockIdx.y = by ;
        std::unique_lock<std::mutex> lk (P.mu) ;
        P.body = &body ; P.blk = &blk ; P.want = block ; P.running = block ; P.gen++ ;
        P.go.notify_all () ;
        P.done.wait (lk, [&] { return P.running == 0 ; }) ;
    }
    g_block = nullptr ;
}
// A kernel without any block- or warp-level synchronisation (tools/emu_library.py decides that from the
// source text) needs no concurrency: its CUDA threads run one after another in the calling thread.
static inline void launch_seq (dim3 grid, unsigned block, const std::function<void ()> &body)
{
    static std::mutex one_launch ;
    std::lock_guard<std::mutex> serial (one_launch) ;
    gridDim = grid ; blockDim.x = block ;
    g_block = nullptr ; t_warp = nullptr ;
    for (unsigned by = 0 ; by < grid.y ; by++)
    for (unsigned b = 0 ; b < grid.x ; b++)
    {
        blockIdx.x = b ; blockIdx.y = by ;
        for (unsigned t = 0 ; t < block ; t++) { threadIdx.x = t ; t_lane = (int) (t & 31) ; body () ; }
    }
    threadIdx.x = 0 ;
}
// the forms a rewritten `kernel <<<grid, block [, smem [, stream]]>>> (args)` takes
static inline dim3 grid_of (dim3 g) { return g ; }
template <class G> static inline dim3 grid_of (G g) { return dim3 ((unsigned) g) ; }
template <class G, class B> static inline void launch_cfg (G g, B b, const std::function<void ()> &body)
{ launch (grid_of (g), (unsigned) b, body) ; }
template <class G, class B, class S> static inline void launch_cfg (G g, B b, S, const std::function<void ()> &body)
{ launch (grid_of (g), (unsigned) b, body) ; }
template <class G, class B, class S, class T> static inline void launch_cfg (G g, B b, S, T, const std::function<void ()> &body)
{ launch (grid_of (g), (unsigned) b, body) ; }
template <class G, class B> static inline void launch_cfg_seq (G g, B b, const std::function<void ()> &body)
{ launch_seq (grid_of (g), (unsigned) b, body) ; }
template <class G, class B, class S> static inline void launch_cfg_seq (G g, B b, S, const std::function<void ()> &body)
{ launch_seq (grid_of (g), (unsigned) b, body) ; }
template <class G, class B, class S, class T> static inline void launch_cfg_seq (G g, B b, S, T, const std::function<void ()> &body)
{ launch_seq (grid_of (g), (unsigned) b, body) ; }
} // namespace emu

// ---- the slice of the runtime API the library uses: "device" memory is host memory --------------
typedef int cudaError_t ;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInsufficientDriver = 35, cudaErrorNoDevice = 100 } ;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault } ;
enum { cudaStreamNonBlocking = 1 } ;
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 } ;
struct cudaDeviceProp { char name [256] ; int major, minor, multiProcessorCount ; size_t totalGlobalMem ; } ;
struct emuEvent { double ms ; } ;
typedef emuEvent *cudaEvent_t ;
static inline double emu_now_ms ()
{ struct timespec ts ; clock_gettime (CLOCK_MONOTONIC, &ts) ; return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6 ; }
// CUDA IPC has no host stand-in: the peer-memory exchange (engine_peer.cu) reports an error here
struct cudaIpcMemHandle_t { char reserved [64] ; } ;
enum { cudaIpcMemLazyEnablePeerAccess = 1 } ;
static inline cudaError_t cudaIpcGetMemHandle (cudaIpcMemHandle_t *, void *) { return 999 ; }
static inline cudaError_t cudaIpcOpenMemHandle (void **, cudaIpcMemHandle_t, unsigned) { return 999 ; }
static inline cudaError_t cudaIpcCloseMemHandle (void *) { return cudaSuccess ; }
static inline cudaError_t cudaGetDeviceCount (int *n) { *n = 1 ; return cudaSuccess ; }
static inline cudaError_t cudaSetDevice (int) { return cudaSuccess ; }
static inline cudaError_t cudaGetDeviceProperties (cudaDeviceProp *p, int)
{
    memset (p, 0, sizeof (*p)) ; strcpy (p->name, "host emulation (tests/emu)") ;
    p->major = 10 ; p->minor = 0 ; p->multiProcessorCount = 1 ; p->totalGlobalMem = (size_t) 8 << 30 ;
    return cudaSuccess ;
}
static inline cudaError_t cudaMemGetInfo (size_t *f, size_t *t) { *f = (size_t) 4 << 30 ; *t = (size_t) 8 << 30 ; return cudaSuccess ; }
static inline cudaError_t cudaGetLastError () { return cudaSuccess ; }
static inline const char *cudaGetErrorString (cudaError_t) { return "host emulation" ; }
template <class P> static inline cudaError_t cudaMalloc (P **p, size_t n)
{ *p = (P *) aligned_alloc (256, (n + 255) / 256 * 256 + 256) ; return *p ? cudaSuccess : cudaErrorMemoryAllocation ; }
template <class P> static inline cudaError_t cudaMallocAsync (P **p, size_t n, cudaStream_t) { return cudaMalloc (p, n) ; }
template <class P> static inline cudaError_t cudaMallocHost (P **p, size_t n) { return cudaMalloc (p, n) ; }
static inline cudaError_t cudaFree (void *p) { free (p) ; return cudaSuccess ; }
static inline cudaError_t cudaFreeHost (void *p) { free (p) ; return cudaSuccess ; }
static inline cudaError_t cudaMemcpyAsync (void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr)
{ if (n) memmove (d, s, n) ; return cudaSuccess ; }
static inline cudaError_t cudaMemcpy (void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove (d, s, n) ; return cudaSuccess ; }
static inline cudaError_t cudaMemsetAsync (void *d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset (d, v, n) ; return cudaSuccess ; }
static inline cudaError_t cudaMemset (void *d, int v, size_t n) { if (n) memset (d, v, n) ; return cudaSuccess ; }
static inline cudaError_t cudaStreamCreateWithFlags (cudaStream_t *s, unsigned) { *s = (cudaStream_t) 1 ; return cudaSuccess ; }
static inline cudaError_t cudaStreamSynchronize (cudaStream_t) { return cudaSuccess ; }
static inline cudaError_t cudaStreamDestroy (cudaStream_t) { return cudaSuccess ; }
static inline cudaError_t cudaEventCreate (cudaEvent_t *e) { *e = new emuEvent () ; (*e)->ms = 0 ; return cudaSuccess ; }
enum { cudaEventDisableTiming = 2 } ;
static inline cudaError_t cudaEventCreateWithFlags (cudaEvent_t *e, unsigned) { *e = new emuEvent () ; (*e)->ms = 0 ; return cudaSuccess ; }
static inline cudaError_t cudaStreamWaitEvent (cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess ; }
static inline cudaError_t cudaEventDestroy (cudaEvent_t e) { delete e ; return cudaSuccess ; }
static inline cudaError_t cudaEventRecord (cudaEvent_t e, cudaStream_t = nullptr) { e->ms = emu_now_ms () ; return cudaSuccess ; }
static inline cudaError_t cudaEventSynchronize (cudaEvent_t) { return cudaSuccess ; }
static inline cudaError_t cudaEventElapsedTime (float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float) (b->ms - a->ms) ; return cudaSuccess ; }
template <class F> static inline cudaError_t cudaFuncSetAttribute (F, cudaFuncAttribute, int) { return cudaSuccess ; }
