// engine_peer.cu -- the exchange step of the vector multiplies on the N GPUs of one box (SURVEY.md
// 8e: "Vector pull / SpMV: one exchange step ... all-gather of the updated vector slice each
// iteration"), done by the library itself over NVLink peer memory instead of by NCCL calls from the
// host language:
//
//   every rank owns a block of A's vectors = of w's entries.  A gb200_peerbuf is a dense copy of w
//   (values + a presence byte per entry, double-buffered) on EVERY GPU, mapped into every process
//   with CUDA IPC.  gb200_peerbuf_publish writes the entries of the rank's T straight into all N
//   copies with peer stores (one kernel on the library's stream, no staging, no host
//   synchronisation) and then raises the rank's flag on every GPU; gb200_peerbuf_wait spins on the
//   device until all N flags of the local GPU show the current epoch.  After it the local copy holds
//   the whole of w in HBM, ready to be the dense operand of the next GrB_mxv.
//
// Epochs alternate between the two buffers: a rank publishes epoch e+2 only after its own wait of e+1
// returned, which needs every peer's publish of e+1, which follows that peer's last read of epoch e.
#include <vector>
#include "engine.cuh"
#include "semiring.cuh"

struct gb200_peerbuf_s
{
    int64_t n = 0 ;
    int type_code = 0, tsz = 0, rank = 0, world = 1 ;
    gb200::DevBuf local ;                   // this GPU's copy: [flags | values x2 | presence x2]
    size_t off_vals [2], off_pres [2], bytes = 0 ;
    std::vector<void *> base ;              // base [q]: rank q's copy as seen from this process
    gb200::DevBuf dbase ;                   // the same table on the device
    unsigned int epoch = 0 ;
    bool connected = false ;
} ;

namespace gb200 {

constexpr int PEER_MAX = 16 ;
struct PeerTable { unsigned char *base [PEER_MAX] ; } ;

// entries (i, x) of T -> values [i] = x, presence [i] = tag on every GPU
__global__ void peer_scatter_kernel (PeerTable pt, int world, const int32_t *__restrict__ Ti,
    const unsigned char *__restrict__ Tx, int64_t nnz, int tsz, size_t off_vals, size_t off_pres,
    unsigned char tag)
{
    for (int64_t e = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; e < nnz ;
        e += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t i = Ti [e] ;
        uint64_t v8 = 0 ; uint32_t v4 = 0 ; uint16_t v2 = 0 ; uint8_t v1 = 0 ;
        if (tsz == 8) v8 = ((const uint64_t *) Tx) [e] ;
        else if (tsz == 4) v4 = ((const uint32_t *) Tx) [e] ;
        else if (tsz == 2) v2 = ((const uint16_t *) Tx) [e] ;
        else v1 = Tx [e] ;
        for (int q = 0 ; q < world ; q++)
        {
            unsigned char *b = pt.base [q] ;
            if (tsz == 8) ((uint64_t *) (b + off_vals)) [i] = v8 ;
            else if (tsz == 4) ((uint32_t *) (b + off_vals)) [i] = v4 ;
            else if (tsz == 2) ((uint16_t *) (b + off_vals)) [i] = v2 ;
            else (b + off_vals) [i] = v1 ;
            (b + off_pres) [i] = tag ;
        }
    }
}

// the stores of the kernel before this one are performed; raise this rank's flag on every GPU
__global__ void peer_signal_kernel (PeerTable pt, int world, int rank, unsigned int epoch)
{
    const int q = threadIdx.x ;
    if (q >= world) return ;
    __threadfence_system () ;
    volatile unsigned int *flag = (volatile unsigned int *) pt.base [q] + rank ;
    *flag = epoch ;
    __threadfence_system () ;
}

__global__ void peer_wait_kernel (const unsigned int *flags, int world, unsigned int epoch)
{
    const int q = threadIdx.x ;
    if (q < world)
    {
        volatile const unsigned int *f = (volatile const unsigned int *) flags + q ;
        while ((int) (*f - epoch) < 0) { __nanosleep (100) ; }
        __threadfence_system () ;
    }
    __syncthreads () ;
}

// the presence bytes of a buffer are cleared before it is written again
__global__ void peer_clear_kernel (unsigned char *pres, int64_t n)
{
    for (int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x ; t < n ;
        t += (int64_t) gridDim.x * blockDim.x) pres [t] = 0 ;
}

} // namespace gb200

using namespace gb200 ;

extern "C" {
#pragma GCC visibility push(default)

gb200_status gb200_peerbuf_create (gb200_peerbuf *out, int64_t n, int type_code, int rank, int world)
{
    if (out == NULL || n < 0 || world < 1 || world > PEER_MAX || rank < 0 || rank >= world
        || type_code < GB200_BOOL || type_code > GB200_FP64) return GB200_INVALID ;
    *out = NULL ;
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    gb200_peerbuf_s *pb = new (std::nothrow) gb200_peerbuf_s () ;
    if (pb == NULL) return GB200_OUT_OF_MEMORY ;
    pb->n = n ; pb->type_code = type_code ; pb->tsz = type_size (type_code) ;
    pb->rank = rank ; pb->world = world ;
    auto up = [] (size_t v) { return (v + 255) & ~(size_t) 255 ; } ;
    size_t off = up (PEER_MAX * sizeof (unsigned int)) ;
    for (int b = 0 ; b < 2 ; b++) { pb->off_vals [b] = off ; off += up ((size_t) (n > 0 ? n : 1) * pb->tsz) ; }
    for (int b = 0 ; b < 2 ; b++) { pb->off_pres [b] = off ; off += up ((size_t) (n > 0 ? n : 1)) ; }
    pb->bytes = off ;
    gb200_status st = pb->local.alloc (off) ;
    if (st == GB200_SUCCESS && cudaMemsetAsync (pb->local.ptr, 0, off, c.stream) != cudaSuccess) st = GB200_CUDA_ERROR ;
    if (st == GB200_SUCCESS && cudaStreamSynchronize (c.stream) != cudaSuccess) st = GB200_CUDA_ERROR ;
    if (st != GB200_SUCCESS) { delete pb ; return st ; }
    pb->base.assign ((size_t) world, nullptr) ;
    pb->base [(size_t) rank] = pb->local.ptr ;
    *out = pb ;
    return GB200_SUCCESS ;
}

// 64 bytes that another process of this box turns into a mapping of this rank's copy
gb200_status gb200_peerbuf_handle (gb200_peerbuf pb, void *handle64)
{
    if (pb == NULL || handle64 == NULL) return GB200_INVALID ;
    static_assert (sizeof (cudaIpcMemHandle_t) == 64, "CUDA IPC handle size") ;
    cudaIpcMemHandle_t h ;
    GB200_CUDA (cudaIpcGetMemHandle (&h, pb->local.ptr)) ;
    memcpy (handle64, &h, 64) ;
    return GB200_SUCCESS ;
}

// handles: world x 64 bytes in rank order (this rank's own entry is ignored)
gb200_status gb200_peerbuf_connect (gb200_peerbuf pb, const void *handles)
{
    if (pb == NULL || handles == NULL) return GB200_INVALID ;
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    for (int q = 0 ; q < pb->world ; q++)
    {
        if (q == pb->rank) continue ;
        cudaIpcMemHandle_t h ;
        memcpy (&h, (const char *) handles + (size_t) q * 64, 64) ;
        void *p = nullptr ;
        GB200_CUDA (cudaIpcOpenMemHandle (&p, h, cudaIpcMemLazyEnablePeerAccess)) ;
        pb->base [(size_t) q] = p ;
    }
    PeerTable pt ;
    memset (&pt, 0, sizeof (pt)) ;
    for (int q = 0 ; q < pb->world ; q++) pt.base [q] = (unsigned char *) pb->base [(size_t) q] ;
    GB200_TRY (pb->dbase.alloc (sizeof (pt))) ;
    GB200_CUDA (cudaMemcpyAsync (pb->dbase.ptr, &pt, sizeof (pt), cudaMemcpyHostToDevice, c.stream)) ;
    GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
    pb->connected = true ;
    return GB200_SUCCESS ;
}

// T (n-by-1, this rank's block of w) -> every rank's copy, over NVLink; then this rank's flag
gb200_status gb200_peerbuf_publish (gb200_peerbuf pb, gb200_result r)
{
    if (pb == NULL || r == NULL || !pb->connected) return GB200_INVALID ;
    if (r->info.vdim != 1 || r->info.vlen != pb->n || r->info.type_code != pb->type_code)
    {
        set_error ("gb200_peerbuf_publish: T is not a vector of the buffer's length and type") ;
        return GB200_INVALID ;
    }
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    pb->epoch++ ;
    const int b = (int) (pb->epoch & 1u) ;
    const unsigned char tag = 1 ;
    {
        // The other buffer holds the previous epoch, which this rank has finished reading (this publish
        // follows the multiply that read it on the same stream).  Clear its presence bytes now, before
        // this rank's flag goes up: no peer writes the next epoch into it before it has seen that flag.
        unsigned char *pres = (unsigned char *) pb->local.ptr + pb->off_pres [b ^ 1] ;
        int64_t g = (pb->n + 255) / 256 ;
        const int64_t cap = (int64_t) c.sm_count * 8 ;
        if (g > cap) g = cap ;
        if (g > 0) peer_clear_kernel <<<(unsigned) g, 256, 0, c.stream>>> (pres, pb->n) ;
        count_launch () ;
    }
    PeerTable pt ;
    memset (&pt, 0, sizeof (pt)) ;
    for (int q = 0 ; q < pb->world ; q++) pt.base [q] = (unsigned char *) pb->base [(size_t) q] ;
    const int64_t nnz = r->info.nnz ;
    if (nnz > 0)
    {
        int64_t g = (nnz + 255) / 256 ;
        const int64_t cap = (int64_t) c.sm_count * 16 ;
        if (g > cap) g = cap ;
        peer_scatter_kernel <<<(unsigned) g, 256, 0, c.stream>>> (pt, pb->world, r->i.as<int32_t> (),
            (const unsigned char *) r->x.ptr, nnz, pb->tsz, pb->off_vals [b], pb->off_pres [b], tag) ;
        count_launch () ;
    }
    peer_signal_kernel <<<1, 32, 0, c.stream>>> (pt, pb->world, pb->rank, pb->epoch) ;
    count_launch () ;
    GB200_CUDA (cudaGetLastError ()) ;
    return GB200_SUCCESS ;
}

// returns (on the stream) when every rank has published the current epoch
gb200_status gb200_peerbuf_wait (gb200_peerbuf pb)
{
    if (pb == NULL || !pb->connected) return GB200_INVALID ;
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    peer_wait_kernel <<<1, 32, 0, c.stream>>> ((const unsigned int *) pb->local.ptr, pb->world, pb->epoch) ;
    count_launch () ;
    GB200_CUDA (cudaGetLastError ()) ;
    return GB200_SUCCESS ;
}

// the current epoch's whole vector as device pointers (values, presence == *tag where present); the
// stream is synchronised so that the host may hand them to other libraries
gb200_status gb200_peerbuf_view (gb200_peerbuf pb, void **values, unsigned char **presence, int *tag)
{
    if (pb == NULL) return GB200_INVALID ;
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
    const int b = (int) (pb->epoch & 1u) ;
    if (values) *values = (char *) pb->local.ptr + pb->off_vals [b] ;
    if (presence) *presence = (unsigned char *) pb->local.ptr + pb->off_pres [b] ;
    if (tag) *tag = 1 ;
    return GB200_SUCCESS ;
}

// the current epoch's whole vector copied out (host or device destinations); for tests and callers
// that want w outside HBM
gb200_status gb200_peerbuf_read (gb200_peerbuf pb, void *values, unsigned char *presence)
{
    if (pb == NULL) return GB200_INVALID ;
    GB200_TRY (ensure_init ()) ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    const int b = (int) (pb->epoch & 1u) ;
    if (values != NULL && pb->n > 0)
        GB200_CUDA (cudaMemcpyAsync (values, (char *) pb->local.ptr + pb->off_vals [b], (size_t) pb->n * pb->tsz,
            cudaMemcpyDefault, c.stream)) ;
    if (presence != NULL && pb->n > 0)
        GB200_CUDA (cudaMemcpyAsync (presence, (char *) pb->local.ptr + pb->off_pres [b], (size_t) pb->n,
            cudaMemcpyDefault, c.stream)) ;
    GB200_CUDA (cudaStreamSynchronize (c.stream)) ;
    return GB200_SUCCESS ;
}

gb200_status gb200_peerbuf_free (gb200_peerbuf *ppb)
{
    if (ppb == NULL || *ppb == NULL) return GB200_SUCCESS ;
    gb200_peerbuf_s *pb = *ppb ;
    Ctx &c = ctx () ;
    std::lock_guard<std::recursive_mutex> lock (c.mu) ;
    if (c.stream != nullptr) cudaStreamSynchronize (c.stream) ;
    for (int q = 0 ; q < pb->world ; q++)
        if (q != pb->rank && pb->base [(size_t) q] != nullptr) cudaIpcCloseMemHandle (pb->base [(size_t) q]) ;
    cudaGetLastError () ;
    delete pb ;
    *ppb = NULL ;
    return GB200_SUCCESS ;
}

#pragma GCC visibility pop
} // extern "C"
