/* gb_axb_parallel_shim.c -- the reference-side binding of libgb_b200.so.
 *
 * This file is what a maintainer of SuiteSparse:GraphBLAS v2.3.3 would add to route the masked
 * semiring multiply to the B200: it defines GB_AxB_parallel with the exact signature of the
 * reference (Source/GB.h:1522-1537; the original body is Source/GB_AxB_parallel.c:63-154).  In the
 * reference's shared library every internal call goes through the PLT, so a library that exports
 * this symbol and is placed ahead of it in symbol-lookup order (LD_PRELOAD, link order, or
 * dlopen(RTLD_GLOBAL) before the reference) takes over the path under an UNMODIFIED GrB_mxm /
 * GrB_mxv / GrB_vxm caller and an UNMODIFIED reference library.
 *
 * It is compiled against the reference's internal header "GB.h" (it must see struct
 * GB_Matrix_opaque to read M, A, B and to build T); no reference source is copied here.  T is
 * created with the reference's own GB_create so that its arrays come from the allocator the
 * reference will later free or transplant them with (Source/GB_mxm.c:141-159).
 *
 * The compute happens in libgb_b200.so (include/gb_b200.h).  Nothing here computes on the CPU.
 * Semirings outside the 960 built-in workers (user-defined operators are host function
 * pointers) are DECLINED: by default the call fails loudly with GrB_PANIC; with
 * GB200_SHIM_FORWARD=1 it is delegated to the host library's own GB_AxB_parallel (found with
 * dlsym(RTLD_NEXT)) and counted in gb200_shim_stats so that tests can assert it never happened.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "GB.h"
#include "gb_b200.h"

typedef GrB_Info (*axb_parallel_fn) (GrB_Matrix *, GrB_Matrix, const bool, const GrB_Matrix,
    const GrB_Matrix, const GrB_Semiring, const bool, const bool, const GrB_Desc_Value,
    GrB_Desc_Value *, bool *, GB_Context) ;

static int64_t g_gpu_calls = 0, g_forwarded = 0, g_declined = 0 ;
static int g_enabled = -1 ;         /* -1: read GB200_SHIM_DISABLE on first use */
static double g_last_device_ms = 0 ;
static int64_t g_last_flops = 0 ;

__attribute__ ((visibility ("default")))
void gb200_shim_enable (int on) { g_enabled = on ? 1 : 0 ; }

__attribute__ ((visibility ("default")))
void gb200_shim_stats (int64_t *gpu_calls, int64_t *forwarded, int64_t *declined)
{
    if (gpu_calls) *gpu_calls = g_gpu_calls ;
    if (forwarded) *forwarded = g_forwarded ;
    if (declined) *declined = g_declined ;
}

/* GB200_SHIM_STATS=1: one line on stderr when the process ends, so that a test can see that an
 * unmodified program (e.g. the reference's own tri_demo under LD_PRELOAD) really ran on the GPU */
__attribute__ ((destructor))
static void shim_report (void)
{
    if (getenv ("GB200_SHIM_STATS") != NULL)
        fprintf (stderr, "[gb_b200 shim] gpu_calls=%lld forwarded=%lld declined=%lld\n",
            (long long) g_gpu_calls, (long long) g_forwarded, (long long) g_declined) ;
}

__attribute__ ((visibility ("default")))
void gb200_shim_cache (int on, int64_t *hits, int64_t *misses, int64_t *invalidations)
{
    if (on >= 0) gb200_cache_enable (on) ;
    gb200_cache_stats (hits, misses, invalidations, NULL) ;
}

__attribute__ ((visibility ("default")))
void gb200_shim_last (double *device_ms, int64_t *flops)
{
    if (device_ms) *device_ms = g_last_device_ms ;
    if (flops) *flops = g_last_flops ;
}

/* The host library's object-model entry points are looked up at run time (not linked) so that the
 * shim can be loaded before the host library, which is what gives it precedence. */
typedef GrB_Info (*gb_create_fn) (GrB_Matrix *, const GrB_Type, const int64_t, const int64_t,
    const GB_Ap_code, const bool, const int, const double, const int64_t, const int64_t, const bool,
    GB_Context) ;                                           /* Source/GB.h:1021-1035 */
typedef GrB_Info (*gb_free_fn) (GrB_Matrix *) ;             /* Source/GB.h:1125 */
static gb_create_fn host_create = NULL ;
static gb_free_fn host_free = NULL ;

static int bind_host (void)
{
    if (host_create == NULL) host_create = (gb_create_fn) dlsym (RTLD_DEFAULT, "GB_create") ;
    if (host_free == NULL) host_free = (gb_free_fn) dlsym (RTLD_DEFAULT, "GB_free") ;
    return (host_create != NULL && host_free != NULL) ;
}

static axb_parallel_fn host_original (void)
{
    static axb_parallel_fn fn = NULL ;
    if (fn != NULL) return fn ;
    /* LD_PRELOAD / link-order interposition: the host library is next in the search order */
    fn = (axb_parallel_fn) dlsym (RTLD_NEXT, "GB_AxB_parallel") ;
    if (fn == NULL)
    {
        /* dlopen(RTLD_GLOBAL) interposition: find the library that owns GB_AxB_meta (the only
         * caller of this function) and ask it for its own definition */
        void *meta = dlsym (RTLD_DEFAULT, "GB_AxB_meta") ;
        Dl_info di ;
        if (meta != NULL && dladdr (meta, &di) != 0 && di.dli_fname != NULL)
        {
            void *h = dlopen (di.dli_fname, RTLD_LAZY | RTLD_NOLOAD) ;
            if (h != NULL) fn = (axb_parallel_fn) dlsym (h, "GB_AxB_parallel") ;
        }
    }
    if (fn == (axb_parallel_fn) GB_AxB_parallel) fn = NULL ;
    return fn ;
}

/* view a reference matrix through the C ABI; `zero` is a 1-entry fallback for a missing p */
static int as_abi (gb200_matrix *out, const GrB_Matrix A, int64_t **tmp_p)
{
    *tmp_p = NULL ;
    out->vlen = A->vlen ;
    out->vdim = A->vdim ;
    out->nvec = A->nvec ;
    out->h = A->is_hyper ? A->h : NULL ;
    out->i = A->i ;
    out->x = A->x ;
    out->type_code = A->type->code ;
    out->reserved = 0 ;
    if (A->p == NULL || A->nzmax == 0)
    {
        /* GB_NNZ (A) == 0; A->p might not be allocated (Source/GB.h:274-278) */
        *tmp_p = calloc ((size_t) A->nvec + 1, sizeof (int64_t)) ;
        if (*tmp_p == NULL) return (0) ;
        out->p = *tmp_p ;
        out->i = NULL ; out->x = NULL ;
    }
    else out->p = A->p ;
    if (A->is_hyper && A->h == NULL) { out->nvec = 0 ; }
    return (1) ;
}

/* The device copy of a freshly fetched T becomes the resident copy of T's host arrays (a no-op that
 * frees it unless the residency cache is on): the call that follows on the same object -- GrB_reduce
 * after the triangle-counting multiply, the next multiply of a k-truss loop -- then starts from HBM.
 * GB_transplant moves T's arrays into the user's C without copying, so the key survives it. */
static void adopt_result (gb200_result *r, const GrB_Matrix T, const gb200_result_info *f)
{
    gb200_matrix v ;
    v.vlen = f->vlen ; v.vdim = f->vdim ; v.nvec = f->nvec ;
    v.p = T->p ; v.h = f->is_hyper ? T->h : NULL ; v.i = T->i ; v.x = T->x ;
    v.type_code = f->type_code ; v.reserved = 0 ;
    gb200_result_adopt (r, &v) ;
}

__attribute__ ((visibility ("default")))
GrB_Info GB_AxB_parallel            /* same contract as reference Source/GB.h:1522-1537 */
(
    GrB_Matrix *Chandle,
    GrB_Matrix M,
    const bool Mask_comp,
    const GrB_Matrix A,
    const GrB_Matrix B,
    const GrB_Semiring semiring,
    const bool flipxy,
    const bool do_adotb,
    const GrB_Desc_Value AxB_method,
    GrB_Desc_Value *AxB_method_used,
    bool *mask_applied,
    GB_Context Context
)
{
    if (g_enabled < 0) g_enabled = (getenv ("GB200_SHIM_DISABLE") != NULL) ? 0 : 1 ;
    if (!g_enabled)
    {
        /* switched off (used by the parity tests to run the reference in the same process) */
        axb_parallel_fn fn = host_original () ;
        if (fn == NULL) return (GrB_PANIC) ;
        return (fn (Chandle, M, Mask_comp, A, B, semiring, flipxy, do_adotb, AxB_method,
            AxB_method_used, mask_applied, Context)) ;
    }

    (*Chandle) = NULL ;
    GrB_BinaryOp add = semiring->add->op ;
    GrB_BinaryOp mult = semiring->multiply ;

    /* Built-in operators over built-in types (the opcode / type-code part of the test of
     * Source/GB_semiring_builtin.c:59-65).  Operands whose built-in type differs from the multiply
     * operator's input type are accepted: the library casts them on the device exactly as the
     * reference's typecasting path does (GB_AxB_Gustavson.c:360-404, GB_CAST Source/GB.h:2925-2947). */
    bool builtin = !((A->type->code >= GB_UCT_code) || (B->type->code >= GB_UCT_code) ||
        (mult->xtype->code >= GB_UCT_code) || (mult->xtype != mult->ytype) ||
        (add->opcode >= GB_USER_C_opcode) || (mult->opcode >= GB_USER_C_opcode)) ;

    gb200_semiring s ;
    s.add_opcode = add->opcode ;
    s.mult_opcode = mult->opcode ;
    s.xy_code = mult->xtype->code ;
    s.z_code = mult->ztype->code ;
    s.flipxy = flipxy ? 1 : 0 ;

    gb200_result r = NULL ;
    gb200_status st = GB200_NOT_SUPPORTED ;
    int64_t *tp_m = NULL, *tp_a = NULL, *tp_b = NULL ;
    if (builtin)
    {
        gb200_matrix am, bm, mm ;
        int ok = as_abi (&am, A, &tp_a) && as_abi (&bm, B, &tp_b) ;
        if (ok && M != NULL) ok = as_abi (&mm, M, &tp_m) ;
        if (!ok) st = GB200_OUT_OF_MEMORY ;
        else st = gb200_AxB_host (&r, (M != NULL) ? &mm : NULL, Mask_comp ? 1 : 0, &am, &bm, &s,
            do_adotb ? 1 : 0, (int) AxB_method) ;
        free (tp_m) ; free (tp_a) ; free (tp_b) ;
    }

    if (st == GB200_NOT_SUPPORTED)
    {
        __atomic_fetch_add (&g_declined, 1, __ATOMIC_RELAXED) ;    /* user threads may be concurrent */
        if (getenv ("GB200_SHIM_FORWARD") != NULL)
        {
            axb_parallel_fn fn = host_original () ;
            if (fn != NULL)
            {
                __atomic_fetch_add (&g_forwarded, 1, __ATOMIC_RELAXED) ;
                return (fn (Chandle, M, Mask_comp, A, B, semiring, flipxy, do_adotb, AxB_method,
                    AxB_method_used, mask_applied, Context)) ;
            }
        }
        fprintf (stderr, "[gb_b200 shim] GB_AxB_parallel declined (%s); set GB200_SHIM_FORWARD=1 to "
            "delegate such calls to the host library\n", builtin ? gb200_last_error () :
            "semiring or operand types outside the built-in space") ;
        return (GrB_PANIC) ;
    }
    if (st == GB200_OUT_OF_MEMORY) return (GrB_OUT_OF_MEMORY) ;
    if (st != GB200_SUCCESS)
    {
        fprintf (stderr, "[gb_b200 shim] GPU multiply failed: %s\n", gb200_last_error ()) ;
        return ((st == GB200_INVALID) ? GrB_INVALID_VALUE : GrB_PANIC) ;
    }

    /* build T with the reference's allocator (postconditions: SURVEY.md 8b) */
    gb200_result_info f ;
    gb200_result_get_info (r, &f) ;
    if (!bind_host ())
    {
        fprintf (stderr, "[gb_b200 shim] host GraphBLAS library (GB_create/GB_free) not found\n") ;
        gb200_result_free (&r) ;
        return (GrB_PANIC) ;
    }
    GrB_Type ctype = add->ztype ;
    int64_t plen = (f.nvec > 0) ? f.nvec : 1 ;
    GrB_Info info = host_create (Chandle, ctype, f.vlen, f.vdim, GB_Ap_malloc, true,
        GB_SAME_HYPER_AS (f.is_hyper), B->hyper_ratio, plen, (f.nnz > 0) ? f.nnz : 1, true, Context) ;
    if (info != GrB_SUCCESS)
    {
        gb200_result_free (&r) ;
        (*Chandle) = NULL ;
        return (info) ;
    }
    GrB_Matrix C = (*Chandle) ;
    st = gb200_result_fetch (r, C->p, f.is_hyper ? C->h : NULL, C->i, C->x) ;
    if (st != GB200_SUCCESS)
    {
        gb200_result_free (&r) ;
        host_free (Chandle) ;
        (*Chandle) = NULL ;
        return ((st == GB200_OUT_OF_MEMORY) ? GrB_OUT_OF_MEMORY : GrB_PANIC) ;
    }
    adopt_result (&r, C, &f) ;          /* T stays resident if the host switched the residency cache on */
    if (f.is_hyper) C->nvec = f.nvec ;
    C->nvec_nonempty = f.nvec_nonempty ;
    C->magic = GB_MAGIC ;
    (*AxB_method_used) = (GrB_Desc_Value) f.method_used ;
    (*mask_applied) = (f.mask_applied != 0) ;
    __atomic_fetch_add (&g_gpu_calls, 1, __ATOMIC_RELAXED) ;
    g_last_device_ms = f.device_ms ;
    g_last_flops = f.flops ;
    return (GrB_SUCCESS) ;
}

/* -------------------------------------------------------------------------------------------------
 * Operand residency (include/gb_b200.h, gb200_cache_*): the library keeps device copies of operands
 * keyed on their host arrays.  Frees and reallocations reach it through the gb200_host_* allocator;
 * the reference's IN-PLACE writers are interposed here, the same way as GB_AxB_parallel, and report
 * the arrays of the object they are about to change before the original runs:
 *   GB_setElement       Source/GB.h:1990   writes C->x in place when the entry exists
 *   GB_subassign_kernel Source/GB.h:2048   C(I,J)<M> = accum (C(I,J),A): values and zombies in place
 *   GB_wait             Source/GB.h:1931   assembles pending tuples / deletes zombies (only ever called
 *                                          when there is such work, GB_WAIT)
 *   GxB_Matrix_import_* Include/GraphBLAS.h:5751-5886: arrays handed in by the user may be arrays it
 *                                          got from an export and rewrote
 * ------------------------------------------------------------------------------------------------- */
static void *host_symbol (const char *name, void *self)
{
    void *fn = dlsym (RTLD_NEXT, name) ;
    if (fn == NULL || fn == self)
    {
        void *meta = dlsym (RTLD_DEFAULT, "GB_AxB_meta") ;
        Dl_info di ;
        fn = NULL ;
        if (meta != NULL && dladdr (meta, &di) != 0 && di.dli_fname != NULL)
        {
            void *h = dlopen (di.dli_fname, RTLD_LAZY | RTLD_NOLOAD) ;
            if (h != NULL) fn = dlsym (h, name) ;
        }
    }
    return ((fn == self) ? NULL : fn) ;
}

static void report_write (const GrB_Matrix C)
{
    if (C == NULL) return ;
    gb200_cache_invalidate (C->p) ;
    gb200_cache_invalidate (C->h) ;
    gb200_cache_invalidate (C->i) ;
    gb200_cache_invalidate (C->x) ;
}

__attribute__ ((visibility ("default")))
GrB_Info GB_setElement (GrB_Matrix C, const void *scalar, const GrB_Index row, const GrB_Index col,
    const GB_Type_code scalar_code, GB_Context Context)
{
    typedef GrB_Info (*fn_t) (GrB_Matrix, const void *, const GrB_Index, const GrB_Index,
        const GB_Type_code, GB_Context) ;
    static fn_t fn = NULL ;
    if (fn == NULL) fn = (fn_t) host_symbol ("GB_setElement", (void *) GB_setElement) ;
    if (fn == NULL) return (GrB_PANIC) ;
    report_write (C) ;
    return (fn (C, scalar, row, col, scalar_code, Context)) ;
}

__attribute__ ((visibility ("default")))
GrB_Info GB_subassign_kernel (GrB_Matrix C, bool C_replace, const GrB_Matrix M, const bool Mask_comp,
    const GrB_BinaryOp accum, const GrB_Matrix A, const GrB_Index *I, const int64_t ni,
    const GrB_Index *J, const int64_t nj, const bool scalar_expansion, const void *scalar,
    const GB_Type_code scalar_code, GB_Context Context)
{
    typedef GrB_Info (*fn_t) (GrB_Matrix, bool, const GrB_Matrix, const bool, const GrB_BinaryOp,
        const GrB_Matrix, const GrB_Index *, const int64_t, const GrB_Index *, const int64_t, const bool,
        const void *, const GB_Type_code, GB_Context) ;
    static fn_t fn = NULL ;
    if (fn == NULL) fn = (fn_t) host_symbol ("GB_subassign_kernel", (void *) GB_subassign_kernel) ;
    if (fn == NULL) return (GrB_PANIC) ;
    report_write (C) ;
    return (fn (C, C_replace, M, Mask_comp, accum, A, I, ni, J, nj, scalar_expansion, scalar,
        scalar_code, Context)) ;
}

__attribute__ ((visibility ("default")))
GrB_Info GB_wait (GrB_Matrix A, GB_Context Context)
{
    typedef GrB_Info (*fn_t) (GrB_Matrix, GB_Context) ;
    static fn_t fn = NULL ;
    if (fn == NULL) fn = (fn_t) host_symbol ("GB_wait", (void *) GB_wait) ;
    if (fn == NULL) return (GrB_PANIC) ;
    report_write (A) ;
    return (fn (A, Context)) ;
}

#define GB200_REPORT_ARRAY(pp) { if ((pp) != NULL) gb200_cache_invalidate (*(pp)) ; }

__attribute__ ((visibility ("default")))
GrB_Info GxB_Matrix_import_CSR (GrB_Matrix *A, const GrB_Type type, GrB_Index nrows, GrB_Index ncols,
    GrB_Index nvals, int64_t nonempty, GrB_Index **Ap, GrB_Index **Aj, void **Ax, const GrB_Descriptor desc)
{
    typedef GrB_Info (*fn_t) (GrB_Matrix *, const GrB_Type, GrB_Index, GrB_Index, GrB_Index, int64_t,
        GrB_Index **, GrB_Index **, void **, const GrB_Descriptor) ;
    static fn_t fn = NULL ;
    if (fn == NULL) fn = (fn_t) host_symbol ("GxB_Matrix_import_CSR", (void *) GxB_Matrix_import_CSR) ;
    if (fn == NULL) return (GrB_PANIC) ;
    GB200_REPORT_ARRAY (Ap) ; GB200_REPORT_ARRAY (Aj) ; GB200_REPORT_ARRAY (Ax) ;
    return (fn (A, type, nrows, ncols, nvals, nonempty, Ap, Aj, Ax, desc)) ;
}

__attribute__ ((visibility ("default")))
GrB_Info GxB_Matrix_import_CSC (GrB_Matrix *A, const GrB_Type type, GrB_Index nrows, GrB_Index ncols,
    GrB_Index nvals, int64_t nonempty, GrB_Index **Ap, GrB_Index **Ai, void **Ax, const GrB_Descriptor desc)
{
    typedef GrB_Info (*fn_t) (GrB_Matrix *, const GrB_Type, GrB_Index, GrB_Index, GrB_Index, int64_t,
        GrB_Index **, GrB_Index **, void **, const GrB_Descriptor) ;
    static fn_t fn = NULL ;
    if (fn == NULL) fn = (fn_t) host_symbol ("GxB_Matrix_import_CSC", (void *) GxB_Matrix_import_CSC) ;
    if (fn == NULL) return (GrB_PANIC) ;
    GB200_REPORT_ARRAY (Ap) ; GB200_REPORT_ARRAY (Ai) ; GB200_REPORT_ARRAY (Ax) ;
    return (fn (A, type, nrows, ncols, nvals, nonempty, Ap, Ai, Ax, desc)) ;
}

__attribute__ ((visibility ("default")))
GrB_Info GxB_Matrix_import_HyperCSR (GrB_Matrix *A, const GrB_Type type, GrB_Index nrows, GrB_Index ncols,
    GrB_Index nvals, int64_t nonempty, GrB_Index nvec, GrB_Index **Ah, GrB_Index **Ap, GrB_Index **Aj,
    void **Ax, const GrB_Descriptor desc)
{
    typedef GrB_Info (*fn_t) (GrB_Matrix *, const GrB_Type, GrB_Index, GrB_Index, GrB_Index, int64_t,
        GrB_Index, GrB_Index **, GrB_Index **, GrB_Index **, void **, const GrB_Descriptor) ;
    static fn_t fn = NULL ;
    if (fn == NULL) fn = (fn_t) host_symbol ("GxB_Matrix_import_HyperCSR", (void *) GxB_Matrix_import_HyperCSR) ;
    if (fn == NULL) return (GrB_PANIC) ;
    GB200_REPORT_ARRAY (Ah) ; GB200_REPORT_ARRAY (Ap) ; GB200_REPORT_ARRAY (Aj) ; GB200_REPORT_ARRAY (Ax) ;
    return (fn (A, type, nrows, ncols, nvals, nonempty, nvec, Ah, Ap, Aj, Ax, desc)) ;
}

__attribute__ ((visibility ("default")))
GrB_Info GxB_Matrix_import_HyperCSC (GrB_Matrix *A, const GrB_Type type, GrB_Index nrows, GrB_Index ncols,
    GrB_Index nvals, int64_t nonempty, GrB_Index nvec, GrB_Index **Ah, GrB_Index **Ap, GrB_Index **Ai,
    void **Ax, const GrB_Descriptor desc)
{
    typedef GrB_Info (*fn_t) (GrB_Matrix *, const GrB_Type, GrB_Index, GrB_Index, GrB_Index, int64_t,
        GrB_Index, GrB_Index **, GrB_Index **, GrB_Index **, void **, const GrB_Descriptor) ;
    static fn_t fn = NULL ;
    if (fn == NULL) fn = (fn_t) host_symbol ("GxB_Matrix_import_HyperCSC", (void *) GxB_Matrix_import_HyperCSC) ;
    if (fn == NULL) return (GrB_PANIC) ;
    GB200_REPORT_ARRAY (Ah) ; GB200_REPORT_ARRAY (Ap) ; GB200_REPORT_ARRAY (Ai) ; GB200_REPORT_ARRAY (Ax) ;
    return (fn (A, type, nrows, ncols, nvals, nonempty, nvec, Ah, Ap, Ai, Ax, desc)) ;
}

/* -------------------------------------------------------------------------------------------------
 * The neighbours of the multiply below (GB_select, GB_reduce_to_scalar, GB_transpose, GB_accum_mask) follow
 * the decline rule of the seam: a call the device library does not take (GB200_NOT_SUPPORTED: outside the
 * built-in space it implements -- nothing was done) goes to the reference's own function; a call it took
 * and FAILED on (no device, a CUDA error, out of device memory) fails loudly with GrB_PANIC /
 * GrB_OUT_OF_MEMORY and a message, exactly as GB_AxB_parallel above -- nothing is computed on the host
 * behind the caller's back.  GB200_SHIM_FORWARD=1 turns that failure into a delegation (counted).
 * Returns 1 when the caller is to forward the call to the reference, else 0 with *info set.
 * ------------------------------------------------------------------------------------------------- */
static int64_t g_neighbour_failed = 0, g_neighbour_forwarded = 0 ;

__attribute__ ((visibility ("default")))
void gb200_shim_neighbour_stats (int64_t *failed, int64_t *forwarded)
{
    if (failed) *failed = g_neighbour_failed ;
    if (forwarded) *forwarded = g_neighbour_forwarded ;
}

static int neighbour_forward (const char *who, gb200_status st, GrB_Info *info)
{
    if (st == GB200_NOT_SUPPORTED) return (1) ;
    __atomic_fetch_add (&g_neighbour_failed, 1, __ATOMIC_RELAXED) ;
    if (getenv ("GB200_SHIM_FORWARD") != NULL)
    {
        __atomic_fetch_add (&g_neighbour_forwarded, 1, __ATOMIC_RELAXED) ;
        return (1) ;
    }
    fprintf (stderr, "[gb_b200 shim] %s failed on the device (%s); set GB200_SHIM_FORWARD=1 to delegate such "
        "calls to the host library\n", who, gb200_last_error ()) ;
    (*info) = (st == GB200_OUT_OF_MEMORY) ? GrB_OUT_OF_MEMORY : GrB_PANIC ;
    return (0) ;
}

/* -------------------------------------------------------------------------------------------------
 * GxB_select with a built-in operator: GB_select (reference Source/GB.h:1185-1197,
 * body Source/GB_select.c:30-386) is interposed like GB_AxB_parallel.  Only the clean path is taken
 * here -- built-in operator and type, no pending work, checks passed: T = select (A,k) is computed by
 * libgb_b200.so (gb200_select_host), built with the reference's GB_create, and handed to the
 * reference's own GB_accum_mask exactly as GB_select.c:383 does.  Every other call (user-defined
 * operators, errors to report, the quick-mask case) goes to the reference's own GB_select untouched.
 * ------------------------------------------------------------------------------------------------- */
static int64_t g_select_calls = 0 ;

__attribute__ ((visibility ("default")))
int64_t gb200_shim_select_calls (void) { return (g_select_calls) ; }

typedef GrB_Info (*gb_select_fn) (GrB_Matrix, const bool, const GrB_Matrix, const bool,
    const GrB_BinaryOp, const GxB_SelectOp, const GrB_Matrix, const void *, const bool, GB_Context) ;
typedef GrB_Info (*gb_accum_mask_fn) (GrB_Matrix, const GrB_Matrix, const GrB_Matrix, const GrB_BinaryOp,
    GrB_Matrix *, const bool, const bool, GB_Context) ;
typedef GrB_Info (*gb_compatible_fn) (const GrB_Type, const GrB_Matrix, const GrB_Matrix,
    const GrB_BinaryOp, const GrB_Type, GB_Context) ;

__attribute__ ((visibility ("default")))
GrB_Info GB_select (GrB_Matrix C, const bool C_replace, const GrB_Matrix M, const bool Mask_comp,
    const GrB_BinaryOp accum, const GxB_SelectOp op, const GrB_Matrix A, const void *k,
    const bool A_transpose, GB_Context Context)
{
    static gb_select_fn orig = NULL ;
    static gb_accum_mask_fn accum_mask = NULL ;
    static gb_compatible_fn compatible = NULL ;
    if (orig == NULL) orig = (gb_select_fn) host_symbol ("GB_select", (void *) GB_select) ;
    if (accum_mask == NULL) accum_mask = (gb_accum_mask_fn) dlsym (RTLD_DEFAULT, "GB_accum_mask") ;
    if (compatible == NULL) compatible = (gb_compatible_fn) dlsym (RTLD_DEFAULT, "GB_compatible") ;
    if (orig == NULL) return (GrB_PANIC) ;
    if (g_enabled < 0) g_enabled = (getenv ("GB200_SHIM_DISABLE") != NULL) ? 0 : 1 ;
    int mine = g_enabled && accum_mask != NULL && compatible != NULL && bind_host ()
        && C != NULL && A != NULL && op != NULL && op->magic == GB_MAGIC && C->magic == GB_MAGIC
        && A->magic == GB_MAGIC && (M == NULL || M->magic == GB_MAGIC)
        && (accum == NULL || accum->magic == GB_MAGIC)
        && op->opcode <= GB_NONZERO_opcode && A->type->code < GB_UCT_code
        && !(op->opcode < GB_NONZERO_opcode && k == NULL)
        && !GB_PENDING (A) && !GB_ZOMBIES (A) && !GB_PENDING (M) && !GB_ZOMBIES (M)
        && !(Mask_comp && M == NULL) ;                                  /* the quick-mask return */
    if (mine)
    {
        /* the checks of GB_select.c:63-93: a failure is reported by the reference itself */
        int64_t tnrows = (A_transpose) ? GB_NCOLS (A) : GB_NROWS (A) ;
        int64_t tncols = (A_transpose) ? GB_NROWS (A) : GB_NCOLS (A) ;
        mine = (compatible (C->type, C, M, accum, A->type, Context) == GrB_SUCCESS)
            && GB_NROWS (C) == tnrows && GB_NCOLS (C) == tncols ;
    }
    if (!mine) return (orig (C, C_replace, M, Mask_comp, accum, op, A, k, A_transpose, Context)) ;

    /* CSR / CSC and the transposed case: the operator is flipped instead (GB_select.c:118-170) */
    bool A_csc = (A->is_csc == !A_transpose) ;
    int opcode = op->opcode ;
    int64_t kk = (opcode < GB_NONZERO_opcode) ? (*((const int64_t *) k)) : 0 ;
    if (!A_csc && opcode < GB_NONZERO_opcode)
    {
        kk = -kk ;
        if (opcode == GB_TRIL_opcode) opcode = GB_TRIU_opcode ;
        else if (opcode == GB_TRIU_opcode) opcode = GB_TRIL_opcode ;
    }
    gb200_matrix am ;
    int64_t *tp_a = NULL ;
    gb200_result r = NULL ;
    gb200_status st = GB200_OUT_OF_MEMORY ;
    if (as_abi (&am, A, &tp_a)) st = gb200_select_host (&r, &am, opcode, kk) ;
    free (tp_a) ;
    if (st != GB200_SUCCESS)
    {
        GrB_Info fail = GrB_PANIC ;
        if (!neighbour_forward ("GB_select", st, &fail)) return (fail) ;
        return (orig (C, C_replace, M, Mask_comp, accum, op, A, k, A_transpose, Context)) ;
    }
    gb200_result_info f ;
    gb200_result_get_info (r, &f) ;
    GrB_Matrix T = NULL ;
    GrB_Info info = host_create (&T, A->type, A->vlen, A->vdim, GB_Ap_malloc, A_csc,
        GB_SAME_HYPER_AS (f.is_hyper), A->hyper_ratio, (f.nvec > 0) ? f.nvec : 1,
        (f.nnz > 0) ? f.nnz : 1, true, Context) ;
    if (info != GrB_SUCCESS) { gb200_result_free (&r) ; return (info) ; }
    st = gb200_result_fetch (r, T->p, f.is_hyper ? T->h : NULL, T->i, T->x) ;
    if (st != GB200_SUCCESS)
    {
        gb200_result_free (&r) ;
        host_free (&T) ;
        return ((st == GB200_OUT_OF_MEMORY) ? GrB_OUT_OF_MEMORY : GrB_PANIC) ;
    }
    adopt_result (&r, T, &f) ;
    if (f.is_hyper) T->nvec = f.nvec ;
    T->nvec_nonempty = f.nvec_nonempty ;
    T->magic = GB_MAGIC ;
    __atomic_fetch_add (&g_select_calls, 1, __ATOMIC_RELAXED) ;
    return (accum_mask (C, M, NULL, accum, &T, C_replace, Mask_comp, Context)) ;
}

/* -------------------------------------------------------------------------------------------------
 * GrB_reduce of a matrix to a scalar: GB_reduce_to_scalar (reference
 * Source/GB.h, body Source/GB_reduce_to_scalar.c:22-319) is interposed.  The reduction over the entries
 * runs on the device when the monoid is built-in, A has the monoid's type (no typecast per entry) and no
 * pending tuples or zombies; the scalar is then put into a 1-by-1 matrix of the same type and handed to
 * the reference's OWN GB_reduce_to_scalar, which reduces that single entry (identity (+) s = s) and does
 * everything else itself: the checks, the typecast into c, the accumulator.
 * ------------------------------------------------------------------------------------------------- */
static int64_t g_reduce_calls = 0 ;

__attribute__ ((visibility ("default")))
int64_t gb200_shim_reduce_calls (void) { return (g_reduce_calls) ; }

__attribute__ ((visibility ("default")))
GrB_Info GB_reduce_to_scalar (void *c, const GrB_Type ctype, const GrB_BinaryOp accum,
    const GrB_Monoid reduce, const GrB_Matrix A, GB_Context Context)
{
    typedef GrB_Info (*fn_t) (void *, const GrB_Type, const GrB_BinaryOp, const GrB_Monoid,
        const GrB_Matrix, GB_Context) ;
    static fn_t orig = NULL ;
    if (orig == NULL) orig = (fn_t) host_symbol ("GB_reduce_to_scalar", (void *) GB_reduce_to_scalar) ;
    if (orig == NULL) return (GrB_PANIC) ;
    if (g_enabled < 0) g_enabled = (getenv ("GB200_SHIM_DISABLE") != NULL) ? 0 : 1 ;
    int mine = g_enabled && bind_host () && c != NULL && ctype != NULL && A != NULL && reduce != NULL
        && A->magic == GB_MAGIC && reduce->magic == GB_MAGIC && reduce->op != NULL
        && (accum == NULL || accum->magic == GB_MAGIC)
        && reduce->op->opcode < GB_USER_C_opcode && A->type->code < GB_UCT_code
        && A->type == reduce->op->ztype && !GB_PENDING (A) && !GB_ZOMBIES (A)
        && GB_NNZ (A) >= 4096 ;         /* a short loop over host arrays beats a transfer */
    if (!mine) return (orig (c, ctype, accum, reduce, A, Context)) ;
    gb200_matrix am ;
    int64_t *tp_a = NULL ;
    char s [16] ;
    gb200_status st = GB200_OUT_OF_MEMORY ;
    if (as_abi (&am, A, &tp_a)) st = gb200_reduce_host (&am, reduce->op->opcode, s) ;
    free (tp_a) ;
    if (st != GB200_SUCCESS)
    {
        GrB_Info fail = GrB_PANIC ;
        if (!neighbour_forward ("GB_reduce_to_scalar", st, &fail)) return (fail) ;
        return (orig (c, ctype, accum, reduce, A, Context)) ;
    }
    /* S = [s], 1-by-1, one entry */
    GrB_Matrix S = NULL ;
    GrB_Info info = host_create (&S, A->type, 1, 1, GB_Ap_malloc, true, GB_SAME_HYPER_AS (false),
        A->hyper_ratio, 1, 1, true, Context) ;
    if (info != GrB_SUCCESS) return (orig (c, ctype, accum, reduce, A, Context)) ;
    S->p [0] = 0 ; S->p [1] = 1 ; S->i [0] = 0 ;
    memcpy (S->x, s, A->type->size) ;
    S->nvec_nonempty = 1 ;
    S->magic = GB_MAGIC ;
    info = orig (c, ctype, accum, reduce, S, Context) ;
    host_free (&S) ;
    __atomic_fetch_add (&g_reduce_calls, 1, __ATOMIC_RELAXED) ;
    return (info) ;
}

/* -------------------------------------------------------------------------------------------------
 * C = (ctype) A' (SURVEY.md 8f row f2): GB_transpose (reference Source/GB.h:2153-2162, body
 * Source/GB_transpose.c:38-985) is interposed.  It is what GB_AxB_meta runs IN FRONT of the multiply for
 * a transposed operand or a mask held in the other format (Source/GB_AxB_meta.c:203,247,311,328-337,355),
 * and what GrB_transpose, GB_accum_mask, GB_eWise ... run for their own transposes, and what a change of
 * format (GxB_Matrix_export_CSR of a matrix held by column, GxB_Matrix_Option_set FORMAT) runs in place.
 * Taken here: all three call forms (GB_transpose.c:55-113: C = A' out of place, C = C' in place of the
 * handle, A = A' in place of the header -- the in-place forms compute T from the intact input and only then
 * swap it in, as the reference's own bucket method does, :913-950) of the general case (avlen > 1, avdim > 1, anz > 0, :470) without an operator (or with the identity of A's own
 * type, which the reference drops too, :213-220), built-in types, no pending work.  T comes from the
 * device in the form the reference's method would have produced (quicksort: hypersparse, bucket: not;
 * the memory estimate of :497-606 restated) already conformed by the rule of GB_to_hyper_conform; the
 * reference's own GB_to_hyper_conform is then run on it as :968-975 does (it finds nothing to change).
 * Everything else -- vectors, empty matrices, operators, user-defined types -- goes to the reference's own
 * GB_transpose untouched.
 * ------------------------------------------------------------------------------------------------- */
static int64_t g_transpose_calls = 0 ;
static int64_t g_transpose_min = -1 ;       /* fewer entries than this: the host's own loop is faster */

__attribute__ ((visibility ("default")))
int64_t gb200_shim_transpose_calls (void) { return (g_transpose_calls) ; }

__attribute__ ((visibility ("default")))
void gb200_shim_transpose_min (int64_t nnz) { g_transpose_min = (nnz < 0) ? 0 : nnz ; }

__attribute__ ((visibility ("default")))
GrB_Info GB_transpose (GrB_Matrix *Chandle, GrB_Type ctype, const bool C_is_csc, const GrB_Matrix A_in,
    const GrB_UnaryOp op_in, GB_Context Context)
{
    typedef GrB_Info (*fn_t) (GrB_Matrix *, GrB_Type, const bool, const GrB_Matrix, const GrB_UnaryOp,
        GB_Context) ;
    typedef GrB_Info (*conform_fn) (GrB_Matrix, GB_Context) ;
    typedef GrB_Info (*transplant_fn) (GrB_Matrix, const GrB_Type, GrB_Matrix *, GB_Context) ;
    static fn_t orig = NULL ;
    static conform_fn conform = NULL ;
    static transplant_fn transplant = NULL ;
    if (orig == NULL) orig = (fn_t) host_symbol ("GB_transpose", (void *) GB_transpose) ;
    if (conform == NULL) conform = (conform_fn) dlsym (RTLD_DEFAULT, "GB_to_hyper_conform") ;
    if (transplant == NULL) transplant = (transplant_fn) dlsym (RTLD_DEFAULT, "GB_transplant") ;
    if (orig == NULL) return (GrB_PANIC) ;
    if (g_enabled < 0) g_enabled = (getenv ("GB200_SHIM_DISABLE") != NULL) ? 0 : 1 ;
    if (g_transpose_min < 0)
    {
        const char *env = getenv ("GB200_TRANSPOSE_MIN_NNZ") ;
        g_transpose_min = (env != NULL && atoll (env) >= 0) ? atoll (env) : 65536 ;
    }
    /* the three call forms of GB_transpose.c:55-113 */
    const int in_place_C = (A_in == NULL) ;                                 /* C = C': *Chandle replaced */
    const int in_place_A = (A_in != NULL && (Chandle == NULL || (*Chandle) == A_in)) ;  /* A = A', header kept */
    const GrB_Matrix A = in_place_C ? ((Chandle != NULL) ? (*Chandle) : NULL) : A_in ;
    int mine = g_enabled && conform != NULL && transplant != NULL && bind_host () && A != NULL
        && A->magic == GB_MAGIC && A->type != NULL && A->type->code < GB_UCT_code
        && (ctype == NULL || ctype->code < GB_UCT_code)
        && (op_in == NULL || (op_in->opcode == GB_IDENTITY_opcode && A->type == op_in->xtype))
        && !GB_PENDING (A) && !GB_ZOMBIES (A) && A->vlen > 1 && A->vdim > 1
        && GB_NNZ (A) > 0 && GB_NNZ (A) >= g_transpose_min ;
    if (!mine) return (orig (Chandle, ctype, C_is_csc, A_in, op_in, Context)) ;
    if (ctype == NULL || op_in != NULL) ctype = A->type ;          /* GB_transpose.c:203-220 */
    const double A_hyper_ratio = A->hyper_ratio ;

    /* which of the reference's two methods would run decides the form T starts out in (:482-606) */
    int via_qsort = 1 ;
    if (!A->is_hyper)
    {
        const int64_t anz = GB_NNZ (A), avlen = A->vlen ;
        const size_t csize = ctype->size ;
        const int in_place = in_place_A || in_place_C ;
        const int recycle_Ai = in_place && !A->i_shallow ;
        double qusage = 0, qsort_memory = 0 ;
        if (!recycle_Ai) qusage += GBYTES (anz, sizeof (int64_t)) ;    /* Tj, unless A->i is recycled */
        qusage += GBYTES (anz, sizeof (int64_t)) ;              /* Ti */
        qsort_memory = qusage ;
        if (in_place && !A->p_shallow) qusage -= GBYTES (A->plen+1, sizeof (int64_t)) ;    /* Ap freed early */
        qusage += GBYTES (anz, sizeof (int64_t)) ;              /* kwork of GB_builder */
        qsort_memory = GB_IMAX (qsort_memory, qusage) ;
        qusage -= GBYTES (anz, sizeof (int64_t)) ;              /* Tj freed */
        qusage += GBYTES (anz, csize) ;                         /* T->x */
        qsort_memory = GB_IMAX (qsort_memory, qusage) ;
        double bucket_memory = GBYTES (avlen, sizeof (int64_t)) + GBYTES (anz, sizeof (int64_t))
            + GBYTES (anz, csize) + GBYTES (avlen, sizeof (int64_t)) ;
        via_qsort = (qsort_memory < bucket_memory) ;
    }

    gb200_matrix am ;
    int64_t *tp_a = NULL ;
    gb200_result r = NULL ;
    gb200_status st = GB200_OUT_OF_MEMORY ;
    if (as_abi (&am, A, &tp_a)) st = gb200_transpose_host (&r, &am, ctype->code, via_qsort, A_hyper_ratio) ;
    free (tp_a) ;
    GrB_Info info = GrB_SUCCESS ;
    GrB_Matrix T = NULL ;
    if (st != GB200_SUCCESS)
    {
        GrB_Info fail = GrB_PANIC ;
        if (neighbour_forward ("GB_transpose", st, &fail))
            return (orig (Chandle, ctype, C_is_csc, A_in, op_in, Context)) ;
        info = fail ;
    }
    gb200_result_info f ;
    if (info == GrB_SUCCESS)
    {
        gb200_result_get_info (r, &f) ;
        info = host_create (&T, ctype, f.vlen, f.vdim, GB_Ap_malloc, C_is_csc,
            GB_SAME_HYPER_AS (f.is_hyper), A_hyper_ratio, (f.nvec > 0) ? f.nvec : 1,
            (f.nnz > 0) ? f.nnz : 1, true, Context) ;
    }
    if (info == GrB_SUCCESS)
    {
        st = gb200_result_fetch (r, T->p, f.is_hyper ? T->h : NULL, T->i, T->x) ;
        if (st != GB200_SUCCESS) info = (st == GB200_OUT_OF_MEMORY) ? GrB_OUT_OF_MEMORY : GrB_PANIC ;
    }
    const int64_t *p_fetched = NULL ;
    if (info == GrB_SUCCESS)
    {
        if (f.is_hyper) T->nvec = f.nvec ;
        T->nvec_nonempty = f.nvec_nonempty ;
        T->magic = GB_MAGIC ;
        p_fetched = T->p ;
        info = conform (T, Context) ;                               /* GB_transpose.c:968-975 */
    }
    if (info != GrB_SUCCESS)
    {
        /* the error contract of GB_transpose.c:22-30: a new C or a C transposed in place is freed and
         * returned as NULL; an A transposed in place keeps its header (and, here, its content) */
        if (r != NULL) gb200_result_free (&r) ;
        host_free (&T) ;
        if (in_place_C) host_free (Chandle) ;
        else if (!in_place_A) (*Chandle) = NULL ;
        return (info) ;
    }
    /* the device copy of A' serves the multiply that follows (residency cache on), unless the
     * reference's conform disagreed with the form T arrived in */
    if (T->p == p_fetched && (T->is_hyper ? 1 : 0) == f.is_hyper) adopt_result (&r, T, &f) ;
    else gb200_result_free (&r) ;
    __atomic_fetch_add (&g_transpose_calls, 1, __ATOMIC_RELAXED) ;
    if (in_place_A)
    {
        /* as the bucket method in place of A (:936-944): the content of A is freed, T moved into its header */
        return (transplant (A, ctype, &T, Context)) ;
    }
    if (in_place_C) host_free (Chandle) ;       /* :920-924: the input C goes, header included */
    (*Chandle) = T ;
    return (GrB_SUCCESS) ;
}

/* -------------------------------------------------------------------------------------------------
 * C<M> = accum (C,T): GB_accum_mask (reference Source/GB.h, body
 * Source/GB_accum_mask.c:130-328) is interposed: the step every GrB_mxm / mxv / vxm, GxB_select,
 * GrB_transpose ... ends with.  Taken here: C, T and M (if any) in the same orientation (the reference
 * transposes first otherwise, :173-199 -- such calls go to the reference, whose transposes are interposed
 * above), built-in types, a built-in accumulator or none, no pending tuples or zombies in C, and real work
 * to do (a mask or an accumulator; C = T alone is a transplant, :262-271).  The new C is computed by
 * libgb_b200.so (gb200_accum_mask_host), built with the reference's GB_create in the form GB_mask would give
 * R (hypersparse when C and T both are, GB_mask.c:315), and handed to the reference's own
 * GB_transplant_conform exactly as GB_mask.c ends (:209, final transplant); T is freed as GB_accum_mask
 * does.  Everything else goes to the reference's own GB_accum_mask untouched.
 * ------------------------------------------------------------------------------------------------- */
static int64_t g_accum_mask_calls = 0 ;
/* Fewer entries in C and T together than this: the reference's own host loop.  Unless set (environment
 * GB200_ACCUM_MASK_MIN_NNZ or gb200_shim_accum_mask_min) the device is used from 65536 entries on while the
 * residency cache is on -- T was just adopted by it and C usually is a result adopted earlier, so only the
 * new C crosses PCIe -- and not at all while it is off: T and C would both be uploaded again right after T
 * was fetched, which was not measured against the host's merge this round. */
static int64_t g_accum_mask_min = -1 ;

__attribute__ ((visibility ("default")))
int64_t gb200_shim_accum_mask_calls (void) { return (g_accum_mask_calls) ; }

__attribute__ ((visibility ("default")))
void gb200_shim_accum_mask_min (int64_t nnz) { g_accum_mask_min = nnz ; }    /* < 0: back to the default rule */

__attribute__ ((visibility ("default")))
GrB_Info GB_accum_mask (GrB_Matrix C, const GrB_Matrix M_in, const GrB_Matrix MT_in, const GrB_BinaryOp accum,
    GrB_Matrix *Thandle, const bool C_replace, const bool Mask_complement, GB_Context Context)
{
    typedef GrB_Info (*tc_fn) (GrB_Matrix, GrB_Type, GrB_Matrix *, GB_Context) ;
    static gb_accum_mask_fn orig = NULL ;
    static tc_fn transplant_conform = NULL ;
    if (orig == NULL) orig = (gb_accum_mask_fn) host_symbol ("GB_accum_mask", (void *) GB_accum_mask) ;
    if (transplant_conform == NULL) transplant_conform = (tc_fn) dlsym (RTLD_DEFAULT, "GB_transplant_conform") ;
    if (orig == NULL) return (GrB_PANIC) ;
    if (g_enabled < 0) g_enabled = (getenv ("GB200_SHIM_DISABLE") != NULL) ? 0 : 1 ;
    int64_t min_nnz = g_accum_mask_min ;
    if (min_nnz < 0)
    {
        const char *env = getenv ("GB200_ACCUM_MASK_MIN_NNZ") ;
        if (env != NULL && atoll (env) >= 0) g_accum_mask_min = min_nnz = atoll (env) ;
        else min_nnz = gb200_cache_enabled () ? 65536 : INT64_MAX ;
    }
    GrB_Matrix T = (Thandle != NULL) ? (*Thandle) : NULL ;
    const GrB_Matrix M = M_in ;
    int mine = g_enabled && transplant_conform != NULL && bind_host () && C != NULL && T != NULL
        && C->magic == GB_MAGIC && T->magic == GB_MAGIC && (M == NULL || M->magic == GB_MAGIC)
        && (M != NULL || accum != NULL) && !(M == NULL && Mask_complement)
        && C->is_csc == T->is_csc && (M == NULL || M->is_csc == C->is_csc)
        && C->type->code < GB_UCT_code && T->type->code < GB_UCT_code
        && (M == NULL || M->type->code < GB_UCT_code)
        && (accum == NULL || (accum->magic == GB_MAGIC && accum->opcode >= GB_FIRST_opcode
            && accum->opcode <= GB_LE_opcode && accum->xtype == accum->ytype
            && accum->xtype->code < GB_UCT_code))
        && !GB_PENDING (C) && !GB_ZOMBIES (C) && !GB_PENDING (T) && !GB_ZOMBIES (T)
        && !GB_PENDING (M) && !GB_ZOMBIES (M)
        && C->vlen == T->vlen && C->vdim == T->vdim && C->vdim <= ((int64_t) 1 << 27)
        && GB_NNZ (C) + GB_NNZ (T) >= min_nnz ;
    if (!mine) return (orig (C, M_in, MT_in, accum, Thandle, C_replace, Mask_complement, Context)) ;

    gb200_matrix cm, tm, mm ;
    int64_t *tp_c = NULL, *tp_t = NULL, *tp_m = NULL ;
    gb200_result r = NULL ;
    gb200_status st = GB200_OUT_OF_MEMORY ;
    const int r_hyper = (C->is_hyper && T->is_hyper && C->vdim > 1) ? 1 : 0 ;
    int ok = as_abi (&cm, C, &tp_c) && as_abi (&tm, T, &tp_t) ;
    if (ok && M != NULL) ok = as_abi (&mm, M, &tp_m) ;
    if (ok) st = gb200_accum_mask_host (&r, &cm, &tm, (M != NULL) ? &mm : NULL, Mask_complement ? 1 : 0,
        C_replace ? 1 : 0, (accum != NULL) ? (int) accum->opcode : 0,
        (accum != NULL) ? (int) accum->xtype->code : 0, r_hyper) ;
    free (tp_c) ; free (tp_t) ; free (tp_m) ;
    if (st != GB200_SUCCESS)
    {
        GrB_Info fail = GrB_PANIC ;
        if (!neighbour_forward ("GB_accum_mask", st, &fail))
        {
            host_free (Thandle) ;       /* as GB_accum_mask's own GB_FREE_ALL on an error */
            return (fail) ;
        }
        return (orig (C, M_in, MT_in, accum, Thandle, C_replace, Mask_complement, Context)) ;
    }
    gb200_result_info f ;
    gb200_result_get_info (r, &f) ;
    GrB_Matrix R = NULL ;
    GrB_Info info = host_create (&R, C->type, f.vlen, f.vdim, GB_Ap_malloc, C->is_csc,
        GB_SAME_HYPER_AS (f.is_hyper), C->hyper_ratio, (f.nvec > 0) ? f.nvec : 1,
        (f.nnz > 0) ? f.nnz : 1, true, Context) ;
    if (info == GrB_SUCCESS)
    {
        st = gb200_result_fetch (r, R->p, f.is_hyper ? R->h : NULL, R->i, R->x) ;
        if (st != GB200_SUCCESS)
        {
            host_free (&R) ;
            info = (st == GB200_OUT_OF_MEMORY) ? GrB_OUT_OF_MEMORY : GrB_PANIC ;
        }
    }
    if (info != GrB_SUCCESS)
    {
        /* as GB_accum_mask's own GB_FREE_ALL on an error: T is freed, C is left as it was */
        gb200_result_free (&r) ;
        host_free (Thandle) ;
        return (info) ;
    }
    if (f.is_hyper) R->nvec = f.nvec ;
    R->nvec_nonempty = f.nvec_nonempty ;
    R->magic = GB_MAGIC ;
    adopt_result (&r, R, &f) ;                  /* the new C stays resident if the residency cache is on */
    host_free (Thandle) ;                       /* GB_accum_mask.c: T is freed when done */
    __atomic_fetch_add (&g_accum_mask_calls, 1, __ATOMIC_RELAXED) ;
    return (transplant_conform (C, C->type, &R, Context)) ;
}

/* -------------------------------------------------------------------------------------------------
 * C<M> = accum (C, scalar) over all of C (`v<q> = level`, the other call of the BFS
 * loop, Demo/Source/bfs5m.c:74): GB_assign (reference Source/GB.h:2026-2046, body Source/GB_assign.c) is
 * interposed.  Taken here: scalar expansion with Rows = Cols = GrB_ALL (not a row or column assign), a mask
 * that is neither complemented nor transposed and is held like C, built-in types and accumulator, no
 * pending work, the same size policy as GB_accum_mask above.  The new C is computed by libgb_b200.so
 * (gb200_assign_scalar_host: the scalar on the pattern of the mask's true entries, then the accum / mask
 * kernels) and handed to the reference's own GB_transplant_conform.  Everything else -- index lists, matrix
 * operands, complemented masks (a dense result), errors to report -- goes to the reference's own GB_assign.
 * ------------------------------------------------------------------------------------------------- */
static int64_t g_assign_calls = 0 ;

__attribute__ ((visibility ("default")))
int64_t gb200_shim_assign_calls (void) { return (g_assign_calls) ; }

__attribute__ ((visibility ("default")))
GrB_Info GB_assign (GrB_Matrix C, const bool C_replace, const GrB_Matrix M_in, const bool Mask_comp,
    bool M_transpose, const GrB_BinaryOp accum, const GrB_Matrix A_in, bool A_transpose,
    const GrB_Index *Rows, const GrB_Index nRows_in, const GrB_Index *Cols, const GrB_Index nCols_in,
    const bool scalar_expansion, const void *scalar, const GB_Type_code scalar_code, const bool col_assign,
    const bool row_assign, GB_Context Context)
{
    typedef GrB_Info (*fn_t) (GrB_Matrix, const bool, const GrB_Matrix, const bool, bool, const GrB_BinaryOp,
        const GrB_Matrix, bool, const GrB_Index *, const GrB_Index, const GrB_Index *, const GrB_Index,
        const bool, const void *, const GB_Type_code, const bool, const bool, GB_Context) ;
    typedef GrB_Info (*tc_fn) (GrB_Matrix, GrB_Type, GrB_Matrix *, GB_Context) ;
    static fn_t orig = NULL ;
    static tc_fn transplant_conform = NULL ;
    static const GrB_Index **all = NULL ;
    if (orig == NULL) orig = (fn_t) host_symbol ("GB_assign", (void *) GB_assign) ;
    if (transplant_conform == NULL) transplant_conform = (tc_fn) dlsym (RTLD_DEFAULT, "GB_transplant_conform") ;
    if (all == NULL) all = (const GrB_Index **) dlsym (RTLD_DEFAULT, "GrB_ALL") ;
    if (orig == NULL) return (GrB_PANIC) ;
    if (g_enabled < 0) g_enabled = (getenv ("GB200_SHIM_DISABLE") != NULL) ? 0 : 1 ;
    int64_t min_nnz = g_accum_mask_min ;
    if (min_nnz < 0)
    {
        const char *env = getenv ("GB200_ACCUM_MASK_MIN_NNZ") ;
        if (env != NULL && atoll (env) >= 0) g_accum_mask_min = min_nnz = atoll (env) ;
        else min_nnz = gb200_cache_enabled () ? 65536 : INT64_MAX ;
    }
    const GrB_Matrix M = M_in ;
    int mine = g_enabled && transplant_conform != NULL && all != NULL && bind_host ()
        && scalar_expansion && scalar != NULL && A_in == NULL && !col_assign && !row_assign
        && Rows == (*all) && Cols == (*all) && M != NULL && !Mask_comp && !M_transpose
        && C != NULL && C->magic == GB_MAGIC && M->magic == GB_MAGIC && C->is_csc == M->is_csc
        && C->vlen == M->vlen && C->vdim == M->vdim && C->vdim <= ((int64_t) 1 << 27)
        && C->type->code < GB_UCT_code && M->type->code < GB_UCT_code && scalar_code < GB_UCT_code
        && (accum == NULL || (accum->magic == GB_MAGIC && accum->opcode >= GB_FIRST_opcode
            && accum->opcode <= GB_LE_opcode && accum->xtype == accum->ytype
            && accum->xtype->code < GB_UCT_code))
        && !GB_PENDING (C) && !GB_ZOMBIES (C) && !GB_PENDING (M) && !GB_ZOMBIES (M)
        && GB_NNZ (C) + GB_NNZ (M) >= min_nnz ;
    if (!mine) return (orig (C, C_replace, M_in, Mask_comp, M_transpose, accum, A_in, A_transpose, Rows, nRows_in,
        Cols, nCols_in, scalar_expansion, scalar, scalar_code, col_assign, row_assign, Context)) ;

    gb200_matrix cm, mm ;
    int64_t *tp_c = NULL, *tp_m = NULL ;
    gb200_result r = NULL ;
    gb200_status st = GB200_OUT_OF_MEMORY ;
    /* R as GB_mask would build it: hypersparse when C and Z both are; Z = accum (C,T) is when C and T (the mask's
     * pattern) are, Z = T alone when the mask is */
    const int r_hyper = (C->is_hyper && M->is_hyper && C->vdim > 1) ? 1 : 0 ;
    if (as_abi (&cm, C, &tp_c) && as_abi (&mm, M, &tp_m))
        st = gb200_assign_scalar_host (&r, &cm, &mm, C_replace ? 1 : 0, (accum != NULL) ? (int) accum->opcode : 0,
            (accum != NULL) ? (int) accum->xtype->code : 0, scalar, (int) scalar_code, r_hyper) ;
    free (tp_c) ; free (tp_m) ;
    if (st != GB200_SUCCESS)
    {
        GrB_Info fail = GrB_PANIC ;
        if (!neighbour_forward ("GB_assign", st, &fail)) return (fail) ;
        return (orig (C, C_replace, M_in, Mask_comp, M_transpose, accum, A_in, A_transpose, Rows, nRows_in,
            Cols, nCols_in, scalar_expansion, scalar, scalar_code, col_assign, row_assign, Context)) ;
    }
    gb200_result_info f ;
    gb200_result_get_info (r, &f) ;
    GrB_Matrix R = NULL ;
    GrB_Info info = host_create (&R, C->type, f.vlen, f.vdim, GB_Ap_malloc, C->is_csc,
        GB_SAME_HYPER_AS (f.is_hyper), C->hyper_ratio, (f.nvec > 0) ? f.nvec : 1,
        (f.nnz > 0) ? f.nnz : 1, true, Context) ;
    if (info == GrB_SUCCESS)
    {
        st = gb200_result_fetch (r, R->p, f.is_hyper ? R->h : NULL, R->i, R->x) ;
        if (st != GB200_SUCCESS)
        {
            host_free (&R) ;
            info = (st == GB200_OUT_OF_MEMORY) ? GrB_OUT_OF_MEMORY : GrB_PANIC ;
        }
    }
    if (info != GrB_SUCCESS) { gb200_result_free (&r) ; return (info) ; }       /* C is left as it was */
    if (f.is_hyper) R->nvec = f.nvec ;
    R->nvec_nonempty = f.nvec_nonempty ;
    R->magic = GB_MAGIC ;
    adopt_result (&r, R, &f) ;
    __atomic_fetch_add (&g_assign_calls, 1, __ATOMIC_RELAXED) ;
    return (transplant_conform (C, C->type, &R, Context)) ;
}
