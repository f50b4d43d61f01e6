// engine.cuh -- host-side context, device buffers and error plumbing shared by engine_*.cu
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <mutex>
#include <atomic>
#include <cuda_runtime.h>
#include "../../include/gb_b200.h"
#include "common.cuh"

namespace gb200 {

struct Status
{
    gb200_status code ;
    Status (gb200_status c = GB200_SUCCESS) : code (c) { }
    bool ok () const { return code == GB200_SUCCESS ; }
} ;

void set_error (const char *fmt, ...) ;

#define GB200_CUDA(call)                                                                    \
    do {                                                                                    \
        cudaError_t e__ = (call) ;                                                          \
        if (e__ != cudaSuccess)                                                             \
        {                                                                                   \
            gb200::set_error ("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                 \
                cudaGetErrorString (e__)) ;                                                 \
            return (e__ == cudaErrorMemoryAllocation) ? GB200_OUT_OF_MEMORY : GB200_CUDA_ERROR ; \
        }                                                                                   \
    } while (0)

#define GB200_TRY(expr)                                                                     \
    do { gb200_status s__ = (expr) ; if (s__ != GB200_SUCCESS) return s__ ; } while (0)

struct Ctx
{
    bool ready = false ;
    int device = 0 ;
    int sm_count = 148 ;
    cudaStream_t stream = nullptr ;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr ;
    void *pinned = nullptr ;            // small pinned scratch for scalar read-backs
    size_t pinned_bytes = 0 ;
    std::recursive_mutex mu ;           // one multiply at a time (the seam may be entered by
                                        // many user threads, reference Demo/Program/pthread_demo.c)
    std::atomic<int64_t> launches {0} ;
    std::atomic<int64_t> multiplies {0} ;
    // CUDA events bracketing every semiring-templated launch of the current multiply
    std::vector<cudaEvent_t> kev ;
    int kev_used = 0 ;
    // Side streams for kernels of one multiply that are independent of each other (the owner classes
    // and orientations of the masked dot): forked from `stream` and joined back into it, so that one
    // kernel's tail is filled by the next one's blocks.  group_stream != nullptr: launch_typed launches
    // there and leaves the timing events to group_begin / group_end.
    static constexpr int NSIDE = 4 ;
    cudaStream_t side [NSIDE] = { nullptr, nullptr, nullptr, nullptr } ;
    cudaEvent_t side_done [NSIDE] = { nullptr, nullptr, nullptr, nullptr } ;
    cudaEvent_t fork_ev = nullptr ;
    cudaStream_t group_stream = nullptr ;
    int mask_policy = 0 ;               // of the current multiply: 0 reference rule, 1 keep, 2 drop
    int method_request = 0 ;            // of the current multiply: the GxB_AxB_METHOD asked for
} ;

Ctx &ctx () ;
gb200_status ensure_init () ;

inline void count_launch (int n = 1) { ctx ().launches += n ; }

// dispatch.cu: a group of independent semiring launches spread over the side streams; everything
// queued on the main stream before group_begin is visible to them, everything after group_end sees
// their results.  One pair of timing events brackets the group.
gb200_status group_begin () ;
void group_use (int k) ;            // the following launch_typed calls go to side stream k % NSIDE
gb200_status group_end () ;

// Device workspace comes from a caching allocator of the library's own (engine_util.cu): freed
// blocks are kept by size class and handed out again.  Everything this library does runs on ONE
// stream, so a block that is released while kernels that use it are still queued can safely be given
// to later work on the same stream.  (cudaMallocAsync's pool was measured to re-map memory every few
// multiplies when whole operands are allocated and freed per call -- +80 ms on a 56 ms multiply.)
void *dev_pool_alloc (size_t nbytes, size_t *capacity) ;     // nullptr: out of device memory
void dev_pool_free (void *ptr, size_t capacity) ;
void dev_pool_trim () ;                                      // give every cached block back to the driver
void dev_pool_stats (int64_t *mallocs, int64_t *malloc_us) ; // trips to the driver so far

struct DevBuf
{
    void *ptr = nullptr ;
    size_t bytes = 0 ;
    size_t cap = 0 ;
    DevBuf () { }
    DevBuf (const DevBuf &) = delete ;
    DevBuf &operator= (const DevBuf &) = delete ;
    DevBuf (DevBuf &&o) noexcept : ptr (o.ptr), bytes (o.bytes), cap (o.cap)
    { o.ptr = nullptr ; o.bytes = 0 ; o.cap = 0 ; }
    DevBuf &operator= (DevBuf &&o) noexcept
    {
        if (this != &o)
        {
            release () ;
            ptr = o.ptr ; bytes = o.bytes ; cap = o.cap ;
            o.ptr = nullptr ; o.bytes = 0 ; o.cap = 0 ;
        }
        return *this ;
    }
    ~DevBuf () { release () ; }
    gb200_status alloc (size_t nbytes)
    {
        release () ;
        if (nbytes == 0) nbytes = 16 ;
        ptr = dev_pool_alloc (nbytes, &cap) ;
        if (ptr == nullptr)
        {
            set_error ("device allocation of %zu bytes failed", nbytes) ;
            return GB200_OUT_OF_MEMORY ;
        }
        bytes = nbytes ;
        return GB200_SUCCESS ;
    }
    void release ()
    {
        if (ptr) { dev_pool_free (ptr, cap) ; ptr = nullptr ; bytes = 0 ; cap = 0 ; }
    }
    template <class T> T *as () const { return (T *) ptr ; }
} ;

} // namespace gb200

// ---- opaque handle bodies -----------------------------------------------------------------------
struct gb200_dmatrix_s
{
    gb200::DMat v ;                 // device view
    gb200::DevBuf p, h, i, x ;
    int is_hyper_flag ;             // raw A->is_hyper (h != NULL), as used by GB_AxB_alloc.c:49-50
    int iso_known = 0 ;                 // lazily computed: are all stored values equal
    gb200::DevBuf longitems ;       // lazily built segments of the vectors longer than VEC_LONG
    int64_t n_longitems = 0 ;
    bool has_longitems = false ;
    gb200::DevBuf tilerow ;         // lazily built: stored vector holding entry t * SPMV_TILE
    int64_t n_tiles = 0 ;
    bool has_tilerow = false ;
} ;

struct gb200_result_s
{
    gb200_result_info info ;
    gb200::DevBuf p, h, i, x ;      // p int64 [nvec+1], h int64 [nvec], i int32 [nnz], x Z [nnz]
} ;

namespace gb200 {

// engine_util.cu
gb200_status scan_i64 (const int64_t *in, int64_t *out, int64_t n) ;     // out has n+1 entries
gb200_status scan_u8  (const uint8_t *in, int64_t *out, int64_t n) ;
gb200_status scan_i32 (const int32_t *in, int64_t *out, int64_t n) ;     // items < 2^26 each
gb200_status read_i64 (const int64_t *dptr, int64_t *host) ;             // syncs the stream
gb200_status fill_bits (void *dst, int elem_size, uint64_t bits, int64_t n) ;
uint64_t identity_bits (int z_code, int add_opcode, int *acc_size) ;
gb200_status filter_mask (const gb200_dmatrix_s *M, DMat &Mview, DevBuf &Mp2, DevBuf &Mi2) ;

// the semiring-templated launchers (inst_*.cu)
struct LaunchCfg ;
bool launch_typed (int xy_code, int family, int z_code, int add, int mult, const void *args,
    int grid, int block) ;

// engine_saxpy.cu / engine_dot.cu
gb200_status run_saxpy (gb200_result_s *R, const gb200_dmatrix_s *M, int mask_comp,
    const gb200_dmatrix_s *A, const gb200_dmatrix_s *B, const gb200_semiring &s) ;
gb200_status run_dot (gb200_result_s *R, const gb200_dmatrix_s *M, int mask_comp,
    const gb200_dmatrix_s *A, const gb200_dmatrix_s *B, const gb200_semiring &s) ;
// engine_vec.cu: B is n-by-1
bool vec_shape (const gb200_dmatrix_s *A, const gb200_dmatrix_s *B) ;
gb200_status run_dotv (gb200_result_s *R, const gb200_dmatrix_s *M, int mask_comp,
    const gb200_dmatrix_s *A, const gb200_dmatrix_s *B, const gb200_semiring &s) ;
gb200_status run_saxpyv (gb200_result_s *R, const gb200_dmatrix_s *M, int mask_comp,
    const gb200_dmatrix_s *A, const gb200_dmatrix_s *B, const gb200_semiring &s) ;
gb200_status flopcount (const DMat *M, const DMat &A, const DMat &B, DevBuf &flops, DevBuf &cum,
    int64_t *total) ;

gb200_status ensure_iso (gb200_dmatrix_s *d) ;

// engine_transpose.cu: vecof [e] = name of the vector of A that holds entry e
gb200_status launch_vecof (const DMat &A, int32_t *vecof) ;

// engine_cache.cu: operand residency across calls of the host entry point
gb200_status cache_acquire (gb200_dmatrix *out, const gb200_matrix *host, bool *cached) ;
void cache_release (gb200_dmatrix d) ;
bool cache_insert (gb200_dmatrix_s *d, const gb200_matrix *host) ;
gb200_status cast_values (const void *in, int from, int to, int64_t n, DevBuf &out) ;
gb200_status launch_mask_pos (const DMat &B, const DMat &M, int64_t *lpos) ;

// Turn per-source-vector results into the final T.  `cum` (nsrc+1) is the cumulative entry count
// over the source vectors (B's, or M's, stored vectors), names their vector names (nullptr:
// identity).  Takes ownership of Ci/Cx.
gb200_status assemble (gb200_result_s *R, int64_t nsrc, const int64_t *names, bool src_hyper,
    DevBuf &cum, DevBuf &Ci, DevBuf &Cx, int64_t cnz, bool C_is_hyper, int64_t cvlen,
    int64_t cvdim) ;

// acc (acc_size bytes per slot) -> Z values, optionally gathering flagged slots
gb200_status convert_acc (const void *acc, int acc_size, void *z, int z_code, int64_t n) ;

} // namespace gb200
