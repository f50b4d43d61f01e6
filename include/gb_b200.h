/* gb_b200.h -- C ABI of libgb_b200.so: the B200 (sm_100a) implementation of the
 * masked semiring sparse matrix multiply that sits behind GrB_mxm / GrB_mxv / GrB_vxm
 * in SuiteSparse:GraphBLAS v2.3.3.
 *
 * The boundary is the reference's internal funnel
 *
 *     GrB_Info GB_AxB_parallel (GrB_Matrix *Chandle, GrB_Matrix M, bool Mask_comp,
 *         GrB_Matrix A, GrB_Matrix B, GrB_Semiring semiring, bool flipxy, bool do_adotb,
 *         GrB_Desc_Value AxB_method, GrB_Desc_Value *AxB_method_used, bool *mask_applied,
 *         GB_Context Context)            -- reference Source/GB.h:1522-1537,
 *                                           body Source/GB_AxB_parallel.c:63-154
 *
 * restated here with plain pointers and sizes (no GraphBLAS, CUDA or torch types), so it can
 * be bound from C (the shim in graphblas_b200/csrc/shim/), ctypes, cgo...
 * Matrices are "CSC-agnostic" exactly as at that seam: `vdim` sparse vectors of length
 * `vlen`, standard (h == NULL, nvec == vdim) or hypersparse (h lists the nvec vectors that
 * are present), 64-bit indices, indices ascending inside each vector, no duplicates
 * (reference Source/Template/GB_matrix.h:193-208).
 *
 * Every function returns a gb200_status.  No function ever computes on the CPU: if no
 * sm_100-class device is usable the call fails with GB200_NO_DEVICE.
 */
#ifndef GB_B200_H
#define GB_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (mapped by the shim onto GrB_Info, Include/GraphBLAS.h:208-267) ------- */
typedef enum
{
    GB200_SUCCESS        = 0,   /* -> GrB_SUCCESS                                            */
    GB200_OUT_OF_MEMORY  = 1,   /* -> GrB_OUT_OF_MEMORY (device or host allocation failed)   */
    GB200_NOT_SUPPORTED  = 2,   /* the "decline rule": semiring/type/size outside the        */
                                /* built-in space this library implements; nothing was done  */
    GB200_INVALID        = 3,   /* -> GrB_INVALID_VALUE: malformed arguments                 */
    GB200_NO_DEVICE      = 4,   /* -> GrB_PANIC: no usable CUDA device                       */
    GB200_CUDA_ERROR     = 5    /* -> GrB_PANIC: a CUDA call failed (see gb200_last_error)    */
} gb200_status ;

/* ---- type codes: identical numbering to GB_Type_code, reference Source/GB.h:450-466 ---- */
typedef enum
{
    GB200_BOOL = 0, GB200_INT8 = 1, GB200_UINT8 = 2, GB200_INT16 = 3, GB200_UINT16 = 4,
    GB200_INT32 = 5, GB200_UINT32 = 6, GB200_INT64 = 7, GB200_UINT64 = 8,
    GB200_FP32 = 9, GB200_FP64 = 10
} gb200_type_code ;

/* ---- binary operator codes: identical numbering to GB_Opcode, Source/GB.h:479-550 ------- */
typedef enum
{
    GB200_FIRST = 7, GB200_SECOND = 8, GB200_MIN = 9, GB200_MAX = 10, GB200_PLUS = 11,
    GB200_MINUS = 12, GB200_TIMES = 13, GB200_DIV = 14,
    GB200_ISEQ = 15, GB200_ISNE = 16, GB200_ISGT = 17, GB200_ISLT = 18, GB200_ISGE = 19,
    GB200_ISLE = 20,
    GB200_LOR = 21, GB200_LAND = 22, GB200_LXOR = 23,
    GB200_EQ = 24, GB200_NE = 25, GB200_GT = 26, GB200_LT = 27, GB200_GE = 28, GB200_LE = 29
} gb200_opcode ;

/* ---- method codes: identical values to GrB_Desc_Value, Include/GraphBLAS.h:2803-2823 ---- */
typedef enum
{
    GB200_METHOD_DEFAULT   = 0,
    GB200_METHOD_GUSTAVSON = 1001,  /* saxpy; reported when the GPU hash/bitmap saxpy ran    */
    GB200_METHOD_HEAP      = 1002,  /* accepted as a request, runs the same saxpy kernels    */
    GB200_METHOD_DOT       = 1003,
    /* Optional flags OR-ed into a method request (not GrB_Desc_Value codes).  The saxpy path
     * uses a non-complemented mask only when it pays: it is dropped when total_flops <= nnz(M)
     * (reference Source/GB_AxB_sequential.c:88-95).  A caller that runs one multiply as several
     * slices (graphblas_b200/sharded.py, one slice per GPU) makes that decision once on the
     * global counts and passes it down so that every slice returns the same kind of T: */
    GB200_MASK_KEEP        = 0x10000,   /* use M whatever the local flop count says           */
    GB200_MASK_DROP        = 0x20000    /* ignore M (mask_applied = 0)                        */
} gb200_method ;

/* ---- a borrowed, read-only, host-resident sparse matrix (or n-by-1 vector) -------------- */
typedef struct
{
    int64_t vlen ;          /* length of each sparse vector                                  */
    int64_t vdim ;          /* number of vectors the matrix may hold                         */
    int64_t nvec ;          /* vectors present in p/h: == vdim if h == NULL                  */
    const int64_t *p ;      /* size nvec+1, p[0] == 0, nnz = p[nvec]                         */
    const int64_t *h ;      /* size nvec, ascending; NULL for the standard form              */
    const int64_t *i ;      /* size nnz; may be NULL only if nnz == 0                        */
    const void    *x ;      /* size nnz * sizeof(type); may be NULL only if nnz == 0         */
    int32_t type_code ;     /* gb200_type_code of x                                          */
    int32_t reserved ;
} gb200_matrix ;

/* ---- a semiring, canonicalised the way GB_semiring_builtin does it
 *      (reference Source/GB_semiring_builtin.c:19-151) ------------------------------------ */
typedef struct
{
    int32_t add_opcode ;    /* monoid: MIN MAX PLUS TIMES (non-bool z); LOR LAND LXOR EQ     */
    int32_t mult_opcode ;   /* any gb200_opcode; z = mult (x,y)                              */
    int32_t xy_code ;       /* gb200_type_code of both inputs of mult                        */
    int32_t z_code ;        /* gb200_type_code of the result (== xy_code, or BOOL for EQ..LE) */
    int32_t flipxy ;        /* nonzero: z = mult (b,a).  gb200_semiring_canonical folds it   */
                            /* into the opcode where the reference does; only MINUS and DIV  */
                            /* observe it at run time (reference Source/axb.m:25,27)          */
} gb200_semiring ;

/* Canonicalise (boolean renames, flipxy folding).  Returns GB200_NOT_SUPPORTED if the
 * combination is not one of the 960 built-in workers (Source/GB_AxB_Gustavson_builtin.c:147-200). */
gb200_status gb200_semiring_canonical (gb200_semiring *s) ;

/* ---- library / device -------------------------------------------------------------------- */
gb200_status gb200_init (int device) ;       /* device < 0: use env GB200_DEVICE or 0       */
gb200_status gb200_finalize (void) ;         /* frees the device workspace cache, the stream and its   */
                                             /* events; operand / result handles stay the caller's     */
const char  *gb200_last_error (void) ;       /* thread-local text of the last failure        */
const char  *gb200_version (void) ;
int          gb200_device_count (void) ;

/* ---- device-resident operands (the "resident in HBM" timing mode; also what the host
 *      entry point below uses internally).  Handles are opaque. ----------------------------- */
typedef struct gb200_dmatrix_s *gb200_dmatrix ;

gb200_status gb200_upload (gb200_dmatrix *out, const gb200_matrix *host) ;
/* The same for operand arrays that are already in HBM of this GPU -- the NCCL all-gather of a
 * replicated operand over NVLink ("A replicated or all-gathered") instead of N copies over
 * PCIe.  p, h, i, x of `dev` are device pointers in the same 64-bit layout; nnz = p[nvec] (p[0] == 0). */
gb200_status gb200_upload_from_device (gb200_dmatrix *out, const gb200_matrix *dev, int64_t nnz) ;
gb200_status gb200_dmatrix_free (gb200_dmatrix *d) ;

/* the result T of one multiply, resident on the device until fetched or freed */
typedef struct gb200_result_s *gb200_result ;

typedef struct
{
    int64_t vlen, vdim ;    /* vlen = do_adotb ? A->vdim : A->vlen ;  vdim = B->vdim          */
    int64_t nvec ;          /* entries of p minus one (== vdim if not hypersparse)            */
    int64_t nvec_nonempty ; /* exact                                                          */
    int64_t nnz ;
    int32_t is_hyper ;      /* rule of reference Source/GB_AxB_alloc.c:49-50                   */
    int32_t type_code ;     /* == semiring z_code                                             */
    int32_t method_used ;   /* GB200_METHOD_DOT, or GB200_METHOD_GUSTAVSON (HEAP if HEAP was asked for) */
    int32_t mask_applied ;  /* 1 iff M (including its complement flag) was honoured           */
    int64_t flops ;         /* multiply-add pairs performed (reference GB_AxB_flopcount        */
                            /* definition for saxpy; matched index pairs for dot)             */
    double  device_ms ;     /* CUDA-event time of the compute, operands resident              */
    double  kernel_ms ;     /* of which: the semiring kernels (saxpy numeric / dot), CUDA events */
} gb200_result_info ;

/* C<M> = A*B (do_adotb == 0: A->vdim == B->vlen) or C<M> = A'*B (do_adotb != 0:
 * A->vlen == B->vlen, A is NOT materialised transposed), over `semiring`.
 * M may be NULL.  mask_comp as at the seam.  method: a gb200_method request. */
gb200_status gb200_AxB_device
(
    gb200_result *out,
    gb200_dmatrix M, int mask_comp,
    gb200_dmatrix A, gb200_dmatrix B,
    const gb200_semiring *semiring,
    int do_adotb, int method
) ;

gb200_status gb200_result_get_info (gb200_result r, gb200_result_info *info) ;

/* Copy T into caller-owned arrays (sized from gb200_result_info): p[nvec+1], h[nvec] (only if
 * is_hyper, else pass NULL), i[nnz], x[nnz*sizeof(type)].  The shim passes arrays obtained from the
 * reference's own allocator (GB_create), which is how ownership of T stays with the reference.
 * The destinations may also be device pointers of the same GPU (unified addressing): that is how a
 * multi-GPU caller keeps its slice of T in HBM for the NCCL exchange of SURVEY.md 8(e). */
gb200_status gb200_result_fetch (gb200_result r, int64_t *p, int64_t *h, int64_t *i, void *x) ;
gb200_status gb200_result_free (gb200_result *r) ;

/* ---- the host entry point: what GB_AxB_parallel is replaced by ---------------------------
 * Uploads M, A, B, runs gb200_AxB_device, and leaves the result on the device for
 * gb200_result_fetch.  (Two calls so that the caller can allocate T with its own allocator
 * between them.) */
gb200_status gb200_AxB_host
(
    gb200_result *out,
    const gb200_matrix *M, int mask_comp,
    const gb200_matrix *A, const gb200_matrix *B,
    const gb200_semiring *semiring,
    int do_adotb, int method
) ;

/* ---- GxB_select with the built-in operators (reference Source/GB_select.c:219-323:
 * the step that builds L = tril (A,-1), U = triu (A,1) in front of the triangle-counting multiply).
 * Codes are numbered like GB_Select_Opcode (Source/GB.h:629-642).  In the CSC-agnostic terms of the
 * matrix (j = vector, i = index in it): TRIL keeps j - i <= k, TRIU j - i >= k, DIAG j - i == k,
 * OFFDIAG j - i != k, NONZERO x != 0.  T has A's type, dimensions and hypersparsity. */
typedef enum
{
    GB200_SELECT_TRIL = 0, GB200_SELECT_TRIU = 1, GB200_SELECT_DIAG = 2, GB200_SELECT_OFFDIAG = 3,
    GB200_SELECT_NONZERO = 4
} gb200_select_op ;
gb200_status gb200_select_device (gb200_result *out, gb200_dmatrix A, int select_op, int64_t k) ;
gb200_status gb200_select_host   (gb200_result *out, const gb200_matrix *A, int select_op, int64_t k) ;

/* ---- C = (ctype) A' (reference Source/GB_transpose.c:38-985, the general case
 * :470-985 with no operator: what Source/GB_AxB_meta.c:203,247,311,328-337,355 calls in front of the
 * multiply for a transposed operand or a mask held in the other format).  In the CSC-agnostic terms of
 * the matrix: entry i of vector j of A becomes entry j of vector i of C; C has vlen = A's vdim and
 * vdim = A's vlen, ascending indices in every vector, A's values cast to ctype_code (a gb200_type_code
 * of a built-in type) by the rule of GB_cast_array.  result_hyper: 1 = C hypersparse (only its non-empty
 * vectors listed, what GB_builder returns in the quicksort method, GB_transpose.c:847-862), 0 = standard
 * form (the bucket method, GB_transpose_bucket.c).  hyper_ratio >= 0: that form is then conformed by the
 * rule of GB_to_hyper_conform (GB_transpose.c:968-975; Source/GB_to_hyper_test.c, GB_to_nonhyper_test.c)
 * with this hyper_ratio, so that T arrives in its final form; hyper_ratio < 0: left as asked.  Jumbled
 * vectors of A are fine.  A with 2^32-1 or more entries is declined (GB200_NOT_SUPPORTED). */
gb200_status gb200_transpose_device (gb200_result *out, gb200_dmatrix A, int ctype_code, int result_hyper,
    double hyper_ratio) ;
gb200_status gb200_transpose_host   (gb200_result *out, const gb200_matrix *A, int ctype_code, int result_hyper,
    double hyper_ratio) ;

/* ---- C<M> = accum (C,T) (reference Source/GB_accum_mask.c:130-328 -> Source/GB_add.c
 * (Z = accum (C,T)) -> Source/GB_mask.c (C<M> = Z); MATLAB statements Test/GB_spec_accum.m and
 * Test/GB_spec_mask.m:60-90): the step after the multiply of every masked or accumulated call.  C, T and M
 * (NULL: no mask) have the same vlen and vdim and are held in the same orientation.  The result R has C's
 * type:  Z = T cast to C's type when accum_opcode == 0, else Z(i,j) = accum (C(i,j) cast to x, T(i,j) cast
 * to y) cast to C's type where both exist, C(i,j) where only C does, T(i,j) cast to C's type where only T
 * does; accum_opcode is a gb200_opcode and accum_xy_code the gb200_type_code of its two inputs (boolean
 * renames as in gb200_semiring_canonical).  m(i,j) = (M(i,j) exists and its value is nonzero) != mask_comp,
 * or !mask_comp without a mask.  R(i,j) = Z(i,j) where m (no entry if Z has none), else C(i,j), or no
 * entry when c_replace.  result_hyper: 1 = R lists only its non-empty vectors (GB_mask.c:315: when C and Z
 * are both hypersparse).  More than 2^27 vectors are declined (GB200_NOT_SUPPORTED). */
gb200_status gb200_accum_mask_device (gb200_result *out, gb200_dmatrix C, gb200_dmatrix T, gb200_dmatrix M,
    int mask_comp, int c_replace, int accum_opcode, int accum_xy_code, int result_hyper) ;
gb200_status gb200_accum_mask_host (gb200_result *out, const gb200_matrix *C, const gb200_matrix *T,
    const gb200_matrix *M, int mask_comp, int c_replace, int accum_opcode, int accum_xy_code,
    int result_hyper) ;

/* ---- C<M> = accum (C, scalar) over all of C (`v<q> = level`, the other call of the
 * BFS loop, reference Demo/Source/bfs5m.c:74 -> Source/GB_assign_scalar.c -> Source/GB_assign.c with
 * I = J = GrB_ALL and a non-complemented mask).  The expanded scalar is dense, but C<M> = Z only looks at it
 * where the mask admits: T = the scalar (a value of type scalar_code) on the pattern of M's true entries,
 * then exactly gb200_accum_mask_* above with mask_comp = 0.  R has C's type: where M admits, accum (C, scalar)
 * (the scalar alone where C has no entry or there is no accumulator); elsewhere C's entry, or none when
 * c_replace. */
gb200_status gb200_assign_scalar_device (gb200_result *out, gb200_dmatrix C, gb200_dmatrix M, int c_replace,
    int accum_opcode, int accum_xy_code, const void *scalar, int scalar_code, int result_hyper) ;
gb200_status gb200_assign_scalar_host (gb200_result *out, const gb200_matrix *C, const gb200_matrix *M,
    int c_replace, int accum_opcode, int accum_xy_code, const void *scalar, int scalar_code, int result_hyper) ;

/* ---- GrB_reduce of a matrix to a scalar over a built-in monoid (reference
 * Source/GB_reduce_to_scalar.c:107-270).  add_opcode: a gb200_opcode naming the monoid (MIN MAX PLUS
 * TIMES, or LOR LAND LXOR EQ for bool; boolean renames as in gb200_semiring_canonical).  *scalar
 * receives one value of A's type; an empty A gives the monoid's identity.  Entries are taken as they
 * are: the caller deals with zombies, typecasting and the accumulator (the shim does). */
gb200_status gb200_reduce_device (gb200_dmatrix A, int add_opcode, void *scalar) ;
gb200_status gb200_reduce_host   (const gb200_matrix *A, int add_opcode, void *scalar) ;
/* the same over the values of a result still on the device: the checksum of a slab of C that is
 * computed, summed and discarded when the whole C would not fit HBM */
gb200_status gb200_result_reduce (gb200_result r, int add_opcode, void *scalar) ;

/* Hand a result over to the residency cache: call after gb200_result_fetch, instead of
 * gb200_result_free, with the host view (p, h, i, x as fetched) -- the device copy of T then serves the
 * next call on those arrays.  Always consumes *r; frees it when the cache is off or T is not kept. */
gb200_status gb200_result_adopt (gb200_result *r, const gb200_matrix *host) ;

/* ---- operand residency across gb200_AxB_host calls (SURVEY.md 8b "Residency") ----------------
 * With the cache on, gb200_AxB_host keeps the device copy of every matrix operand (not of vectors)
 * keyed on the addresses and shape of its host arrays, so that the graph of a BFS / SSSP / k-truss
 * loop crosses PCIe once.  An entry is dropped when one of its arrays is freed or reallocated
 * (gb200_host_free / gb200_host_realloc report that themselves: use them as the GxB_init allocator),
 * when the caller reports a write (gb200_cache_invalidate: the shim does so for the reference's
 * in-place writers GB_setElement, GB_subassign_kernel, GB_wait and for GxB_*_import_*), or when a
 * sampled fingerprint of the host arrays no longer matches.  Off by default (GB200_OPERAND_CACHE=1 or
 * gb200_cache_enable (1)); bounded to a quarter of the device's memory (GB200_OPERAND_CACHE_MB). */
void gb200_cache_enable (int on) ;
int  gb200_cache_enabled (void) ;
void gb200_cache_invalidate (const void *array) ;    /* any of p, h, i, x of a cached operand       */
void gb200_cache_clear (void) ;
void gb200_cache_stats (int64_t *hits, int64_t *misses, int64_t *invalidations, int64_t *resident_bytes) ;

/* ---- GB_AxB_flopcount on the device (reference Source/GB_AxB_flopcount.c:85-316) --------
 * Bflops_out (host, size B->nvec+1, may be NULL) receives the cumulative sum; *total the last
 * entry. */
gb200_status gb200_flopcount_device
(
    gb200_dmatrix M, gb200_dmatrix A, gb200_dmatrix B,
    int64_t *Bflops_out, int64_t *total
) ;

/* ---- multi-GPU: flop-balanced 1-D partition of B's vectors (the slicing GB_AxB_parallel.c:52
 * and GB_AxB_flopcount.c:32-37 plan for).  Given the cumulative Bflops (size nvec+1) returns
 * nparts+1 vector boundaries in `bounds`. */
gb200_status gb200_partition_by_flops
(
    const int64_t *Bflops_cumulative, int64_t nvec, int nparts, int64_t *bounds
) ;

/* ---- the exchange step of the vector multiplies on the N GPUs of one box ----------
 * One process per GPU; every rank owns a block of A's vectors = of w's entries.  A gb200_peerbuf is a
 * dense copy of w (values + one presence byte per entry) on every GPU, mapped into every process with
 * CUDA IPC.  publish writes the entries of the rank's T (n-by-1) into ALL copies with peer stores over
 * NVLink -- one kernel on the library's stream, no staging, no host synchronisation -- and raises the
 * rank's flag on every GPU; wait returns (on the stream) once every rank has published the same
 * epoch.  The handles (64 bytes each) are exchanged by the host language (torch.distributed, MPI...). */
typedef struct gb200_peerbuf_s *gb200_peerbuf ;
gb200_status gb200_peerbuf_create  (gb200_peerbuf *out, int64_t n, int type_code, int rank, int world) ;
gb200_status gb200_peerbuf_handle  (gb200_peerbuf pb, void *handle64) ;
gb200_status gb200_peerbuf_connect (gb200_peerbuf pb, const void *handles /* world x 64 bytes */) ;
gb200_status gb200_peerbuf_publish (gb200_peerbuf pb, gb200_result T) ;
gb200_status gb200_peerbuf_wait    (gb200_peerbuf pb) ;
/* device pointers of the current epoch's whole vector: values [n], presence [n] (== *tag where present) */
gb200_status gb200_peerbuf_view    (gb200_peerbuf pb, void **values, unsigned char **presence, int *tag) ;
gb200_status gb200_peerbuf_read    (gb200_peerbuf pb, void *values, unsigned char *presence) ;  /* copies out */
gb200_status gb200_peerbuf_free    (gb200_peerbuf *pb) ;

/* ---- pinned host memory ---------------------------------------------------------------------
 * A malloc / calloc / realloc / free quartet with the signatures GxB_init takes (reference
 * Include/GraphBLAS.h:330-340).  A host application that starts the reference with
 *     GxB_init (mode, gb200_host_malloc, gb200_host_calloc, gb200_host_realloc, gb200_host_free)
 * keeps every GraphBLAS array (operands, and the T that the shim builds with GB_create) in
 * page-locked memory, so the copies of gb200_AxB_host / gb200_result_fetch run at PCIe speed.
 * Blocks of at least GB200_HOST_PIN_MIN bytes are page-locked and recycled through a size-class
 * cache (page-locking costs ~0.3 ms per MiB); smaller ones come from malloc.  Usable before
 * gb200_init; without a CUDA device they fall back to plain malloc (memory, not compute). */
#define GB200_HOST_PIN_MIN (1 << 16)
void *gb200_host_malloc  (size_t size) ;
void *gb200_host_calloc  (size_t n, size_t size) ;
void *gb200_host_realloc (void *p, size_t size) ;
void  gb200_host_free    (void *p) ;
void  gb200_host_trim    (void) ;            /* release the cached page-locked blocks       */
/* Both caches are bounded: freed page-locked blocks are kept up to GB200_HOST_CACHE_MB (default: a
 * quarter of the machine's RAM), freed device blocks up to GB200_DEVICE_CACHE_MB (default: half of
 * the device's memory); beyond that the largest free blocks go back to the driver. */
void  gb200_device_trim  (void) ;            /* release the cached free device blocks       */

/* ---- timing on the library's own stream (CUDA events) ----------------------------------------
 * gb200_timer_mark records event `slot` (0..7) on the stream every kernel of this library is
 * launched on; gb200_timer_elapsed_ms synchronises on slot_b and returns the device time between
 * the two marks.  This is how bench.py times K steps on the launching stream. */
gb200_status gb200_timer_mark (int slot) ;
gb200_status gb200_timer_elapsed_ms (int slot_a, int slot_b, double *ms) ;

/* counters: how many kernels this library has launched / how many multiplies it has run */
int64_t gb200_kernel_launches (void) ;
int64_t gb200_multiplies (void) ;

#ifdef __cplusplus
}
#endif
#endif
