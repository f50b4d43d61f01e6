/* gb_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A plain-C CPU restatement of the reference's masked semiring multiply at the GB_AxB_parallel seam
 * (SuiteSparse:GraphBLAS v2.3.3) and of its neighbours GB_transpose and GB_accum_mask, used only as the
 * checker by tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg and the parity legs of the tools
 * behind bench.py's `neighbours` entry.  Nothing in graphblas_b200/ links, loads or calls it.
 *
 * Parity is PINNED: tests/test_oracle.py checks this file against (a) the compiled reference
 * itself (oracle/_ref, built by Makefile.ref from the unmodified sources) on seeded inputs over every
 * method / mask / format combination, (b) committed golden vectors generated from that reference
 * (tests/golden/, script tests/golden/make_golden.py), and (c) the triangle counts printed in the
 * reference's own Demo/Output/tri_demo.out; oracle_transpose and oracle_accum_mask are pinned against the
 * reference's own GB_transpose / GB_accum_mask / GrB_Matrix_assign called directly.
 *
 * Each function cites the reference code it restates.  The algorithms follow the reference's order
 * of operations exactly (ascending k, identity-or-first-product start), so floating-point results
 * are bit-identical to the reference's Gustavson and dot methods, not merely within tolerance.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdbool.h>
#include <math.h>

/* ---- codes: reference Source/GB.h:450-466 (types) and :479-550 (operators) ------------------- */
enum { T_BOOL, T_INT8, T_UINT8, T_INT16, T_UINT16, T_INT32, T_UINT32, T_INT64, T_UINT64, T_FP32, T_FP64 } ;
enum { OP_FIRST = 7, OP_SECOND, OP_MIN, OP_MAX, OP_PLUS, OP_MINUS, OP_TIMES, OP_DIV,
       OP_ISEQ, OP_ISNE, OP_ISGT, OP_ISLT, OP_ISGE, OP_ISLE, OP_LOR, OP_LAND, OP_LXOR,
       OP_EQ, OP_NE, OP_GT, OP_LT, OP_GE, OP_LE } ;

typedef struct
{
    int64_t vlen, vdim, nvec ;
    const int64_t *p, *h, *i ;
    const void *x ;
    int32_t type_code, is_hyper ;       /* is_hyper: h != NULL */
} omat ;

typedef struct
{
    int64_t vlen, vdim, nvec, nvec_nonempty, nnz ;
    int64_t *p, *h, *i ;
    void *x ;
    int32_t is_hyper, type_code, mask_applied, method_used ;
} oresult ;

static const size_t tsize [11] = { 1, 1, 1, 2, 2, 4, 4, 8, 8, 4, 8 } ;

/* ---- integer division, reference Source/GB.h:2782-2870 (GB_IDIV) ----------------------------- */
#define SDIV(T, LO, HI) static inline T sdiv_##T (T x, T y) { \
    if (y == -1) return (T) (0 - (uint64_t) x) ; \
    if (y == 0) return (x == 0) ? 0 : ((x < 0) ? LO : HI) ; return (T) (x / y) ; }
#define UDIV(T, HI) static inline T udiv_##T (T x, T y) { \
    if (y == 0) return (x == 0) ? 0 : HI ; return (T) (x / y) ; }
SDIV (int8_t, INT8_MIN, INT8_MAX) SDIV (int16_t, INT16_MIN, INT16_MAX)
SDIV (int32_t, INT32_MIN, INT32_MAX) SDIV (int64_t, INT64_MIN, INT64_MAX)
UDIV (uint8_t, UINT8_MAX) UDIV (uint16_t, UINT16_MAX) UDIV (uint32_t, UINT32_MAX)
UDIV (uint64_t, UINT64_MAX)

/* ---- multiply operators, reference Source/axb.m:19-45; z = mult (x,y), flip only for MINUS/DIV -- */
#define MULT_NUM(T, NAME, MINF, MAXF, DIVF, UT)                                                  \
static void mult_##NAME (int op, const void *xp, const void *yp, void *zp, int flip)             \
{                                                                                                \
    T x = *(const T *) xp, y = *(const T *) yp ;                                                 \
    switch (op)                                                                                  \
    {                                                                                            \
        case OP_FIRST  : *(T *) zp = x ; break ;                                                 \
        case OP_SECOND : *(T *) zp = y ; break ;                                                 \
        case OP_MIN    : *(T *) zp = MINF ; break ;                                              \
        case OP_MAX    : *(T *) zp = MAXF ; break ;                                              \
        case OP_PLUS   : *(T *) zp = (T) ((UT) x + (UT) y) ; break ;                             \
        case OP_MINUS  : *(T *) zp = flip ? (T) ((UT) y - (UT) x) : (T) ((UT) x - (UT) y) ; break ; \
        case OP_TIMES  : *(T *) zp = (T) ((UT) x * (UT) y) ; break ;                             \
        case OP_DIV    : { T a = flip ? y : x, b = flip ? x : y ; *(T *) zp = DIVF ; } break ;   \
        case OP_ISEQ   : *(T *) zp = (T) (x == y) ; break ;                                      \
        case OP_ISNE   : *(T *) zp = (T) (x != y) ; break ;                                      \
        case OP_ISGT   : *(T *) zp = (T) (x >  y) ; break ;                                      \
        case OP_ISLT   : *(T *) zp = (T) (x <  y) ; break ;                                      \
        case OP_ISGE   : *(T *) zp = (T) (x >= y) ; break ;                                      \
        case OP_ISLE   : *(T *) zp = (T) (x <= y) ; break ;                                      \
        case OP_LOR    : *(T *) zp = (T) ((x != 0) || (y != 0)) ; break ;                        \
        case OP_LAND   : *(T *) zp = (T) ((x != 0) && (y != 0)) ; break ;                        \
        case OP_LXOR   : *(T *) zp = (T) ((x != 0) != (y != 0)) ; break ;                        \
        case OP_EQ     : *(bool *) zp = (x == y) ; break ;                                       \
        case OP_NE     : *(bool *) zp = (x != y) ; break ;                                       \
        case OP_GT     : *(bool *) zp = (x >  y) ; break ;                                       \
        case OP_LT     : *(bool *) zp = (x <  y) ; break ;                                       \
        case OP_GE     : *(bool *) zp = (x >= y) ; break ;                                       \
        case OP_LE     : *(bool *) zp = (x <= y) ; break ;                                       \
    }                                                                                            \
}
#define IMIN ((x < y) ? x : y)
#define IMAX ((x > y) ? x : y)
MULT_NUM (int8_t,   int8,   IMIN, IMAX, sdiv_int8_t (a, b),   uint32_t)
MULT_NUM (uint8_t,  uint8,  IMIN, IMAX, udiv_uint8_t (a, b),  uint32_t)
MULT_NUM (int16_t,  int16,  IMIN, IMAX, sdiv_int16_t (a, b),  uint32_t)
MULT_NUM (uint16_t, uint16, IMIN, IMAX, udiv_uint16_t (a, b), uint32_t)
MULT_NUM (int32_t,  int32,  IMIN, IMAX, sdiv_int32_t (a, b),  uint32_t)
MULT_NUM (uint32_t, uint32, IMIN, IMAX, udiv_uint32_t (a, b), uint32_t)
MULT_NUM (int64_t,  int64,  IMIN, IMAX, sdiv_int64_t (a, b),  uint64_t)
MULT_NUM (uint64_t, uint64, IMIN, IMAX, udiv_uint64_t (a, b), uint64_t)
MULT_NUM (float,    fp32,   fminf (x, y), fmaxf (x, y), (a / b), float)
MULT_NUM (double,   fp64,   fmin (x, y),  fmax (x, y),  (a / b), double)

/* bool: only the operators that survive GB_boolean_rename (Source/GB_boolean_rename.c:36-90) */
static void mult_bool (int op, const void *xp, const void *yp, void *zp, int flip)
{
    bool x = *(const bool *) xp, y = *(const bool *) yp, z ;
    (void) flip ;
    switch (op)
    {
        case OP_FIRST : z = x ; break ;      case OP_SECOND : z = y ; break ;
        case OP_LOR : z = x || y ; break ;   case OP_LAND : z = x && y ; break ;
        case OP_LXOR : z = (x != y) ; break ; case OP_EQ : z = (x == y) ; break ;
        case OP_GT : z = (x > y) ; break ;   case OP_LT : z = (x < y) ; break ;
        case OP_GE : z = (x >= y) ; break ;  default : z = (x <= y) ; break ;
    }
    *(bool *) zp = z ;
}

/* ---- monoids, reference Source/axb_template.m:25-86: w = add (w,t), identity ------------------ */
#define ADD_NUM(T, NAME, MINF, MAXF, UT, IDMIN, IDMAX)                                           \
static void add_##NAME (int op, void *wp, const void *tp)                                        \
{                                                                                                \
    T x = *(T *) wp, y = *(const T *) tp ;                                                       \
    switch (op)                                                                                  \
    {                                                                                            \
        case OP_MIN   : *(T *) wp = MINF ; break ;                                               \
        case OP_MAX   : *(T *) wp = MAXF ; break ;                                               \
        case OP_PLUS  : *(T *) wp = (T) ((UT) x + (UT) y) ; break ;                              \
        default       : *(T *) wp = (T) ((UT) x * (UT) y) ; break ;                              \
    }                                                                                            \
}                                                                                                \
static void ident_##NAME (int op, void *wp)                                                      \
{                                                                                                \
    switch (op)                                                                                  \
    {                                                                                            \
        case OP_MIN   : *(T *) wp = IDMIN ; break ;                                              \
        case OP_MAX   : *(T *) wp = IDMAX ; break ;                                              \
        case OP_PLUS  : *(T *) wp = 0 ; break ;                                                  \
        default       : *(T *) wp = 1 ; break ;                                                  \
    }                                                                                            \
}
ADD_NUM (int8_t,   int8,   IMIN, IMAX, uint32_t, INT8_MAX, INT8_MIN)
ADD_NUM (uint8_t,  uint8,  IMIN, IMAX, uint32_t, UINT8_MAX, 0)
ADD_NUM (int16_t,  int16,  IMIN, IMAX, uint32_t, INT16_MAX, INT16_MIN)
ADD_NUM (uint16_t, uint16, IMIN, IMAX, uint32_t, UINT16_MAX, 0)
ADD_NUM (int32_t,  int32,  IMIN, IMAX, uint32_t, INT32_MAX, INT32_MIN)
ADD_NUM (uint32_t, uint32, IMIN, IMAX, uint32_t, UINT32_MAX, 0)
ADD_NUM (int64_t,  int64,  IMIN, IMAX, uint64_t, INT64_MAX, INT64_MIN)
ADD_NUM (uint64_t, uint64, IMIN, IMAX, uint64_t, UINT64_MAX, 0)
ADD_NUM (float,    fp32,   fminf (x, y), fmaxf (x, y), float, INFINITY, -INFINITY)
ADD_NUM (double,   fp64,   fmin (x, y),  fmax (x, y),  double, INFINITY, -INFINITY)

static void add_bool (int op, void *wp, const void *tp)
{
    bool w = *(bool *) wp, t = *(const bool *) tp ;
    switch (op)
    {
        case OP_LOR : w = (w || t) ; break ;   case OP_LAND : w = (w && t) ; break ;
        case OP_LXOR : w = (w != t) ; break ;  default : w = (w == t) ; break ;     /* EQ */
    }
    *(bool *) wp = w ;
}
static void ident_bool (int op, void *wp)
{
    *(bool *) wp = (op == OP_LAND || op == OP_EQ) ;
}

typedef void (*mult_fn) (int, const void *, const void *, void *, int) ;
typedef void (*add_fn) (int, void *, const void *) ;
typedef void (*ident_fn) (int, void *) ;
static const mult_fn MULT [11] = { mult_bool, mult_int8, mult_uint8, mult_int16, mult_uint16,
    mult_int32, mult_uint32, mult_int64, mult_uint64, mult_fp32, mult_fp64 } ;
static const add_fn ADD [11] = { add_bool, add_int8, add_uint8, add_int16, add_uint16, add_int32,
    add_uint32, add_int64, add_uint64, add_fp32, add_fp64 } ;
static const ident_fn IDENT [11] = { ident_bool, ident_int8, ident_uint8, ident_int16, ident_uint16,
    ident_int32, ident_uint32, ident_int64, ident_uint64, ident_fp32, ident_fp64 } ;

/* ---- the semiring after GB_semiring_builtin's canonicalisation (Source/GB_semiring_builtin.c:86-148) */
typedef struct { int add, mult, xy, z, flip ; } osemiring ;

static int boolean_rename (int op)             /* Source/GB_boolean_rename.c:30-91 */
{
    switch (op)
    {
        case OP_DIV : case OP_FIRST : return OP_FIRST ;
        case OP_MIN : case OP_TIMES : case OP_LAND : return OP_LAND ;
        case OP_MAX : case OP_PLUS : case OP_LOR : return OP_LOR ;
        case OP_MINUS : case OP_ISNE : case OP_NE : case OP_LXOR : return OP_LXOR ;
        case OP_ISEQ : case OP_EQ : return OP_EQ ;
        case OP_ISGT : case OP_GT : return OP_GT ;
        case OP_ISLT : case OP_LT : return OP_LT ;
        case OP_ISGE : case OP_GE : return OP_GE ;
        case OP_ISLE : case OP_LE : return OP_LE ;
        default : return op ;
    }
}

static void canonical (osemiring *s)
{
    if (s->xy == T_BOOL) s->mult = boolean_rename (s->mult) ;
    if (s->z == T_BOOL) s->add = boolean_rename (s->add) ;
    if (s->flip)
    {
        switch (s->mult)
        {
            case OP_FIRST : s->mult = OP_SECOND ; break ;  case OP_SECOND : s->mult = OP_FIRST ; break ;
            case OP_GT : s->mult = OP_LT ; break ;         case OP_LT : s->mult = OP_GT ; break ;
            case OP_GE : s->mult = OP_LE ; break ;         case OP_LE : s->mult = OP_GE ; break ;
            case OP_ISGT : s->mult = OP_ISLT ; break ;     case OP_ISLT : s->mult = OP_ISGT ; break ;
            case OP_ISGE : s->mult = OP_ISLE ; break ;     case OP_ISLE : s->mult = OP_ISGE ; break ;
            default : break ;
        }
    }
}

/* ---- vector lookup: the role of GB_lookup, Source/GB.h:3396-3445 ----------------------------- */
static inline bool IS_HYPER (const omat *A) { return A->is_hyper && A->nvec < A->vdim ; }

static bool lookup (const omat *A, int64_t k, int64_t *pa, int64_t *pe)
{
    if (!IS_HYPER (A)) { *pa = A->p [k] ; *pe = A->p [k+1] ; return *pe > *pa ; }
    int64_t lo = 0, hi = A->nvec - 1 ;
    while (lo <= hi)
    {
        int64_t mid = (lo + hi) / 2 ;
        if (A->h [mid] == k) { *pa = A->p [mid] ; *pe = A->p [mid+1] ; return *pe > *pa ; }
        if (A->h [mid] < k) lo = mid + 1 ; else hi = mid - 1 ;
    }
    *pa = *pe = 0 ;
    return false ;
}
static inline int64_t vecname (const omat *A, int64_t kk) { return IS_HYPER (A) ? A->h [kk] : kk ; }

/* cast_M: a mask entry is true iff its value is nonzero (Gustavson_mask.c:190-197) */
static bool mask_true (const omat *M, int64_t p)
{
    switch (M->type_code)
    {
        case T_BOOL : case T_INT8 : case T_UINT8 : return ((const uint8_t *) M->x) [p] != 0 ;
        case T_INT16 : case T_UINT16 : return ((const uint16_t *) M->x) [p] != 0 ;
        case T_INT32 : case T_UINT32 : return ((const uint32_t *) M->x) [p] != 0 ;
        case T_INT64 : case T_UINT64 : return ((const uint64_t *) M->x) [p] != 0 ;
        case T_FP32 : return ((const float *) M->x) [p] != 0 ;
        default : return ((const double *) M->x) [p] != 0 ;
    }
}

/* ---- GB_AxB_flopcount, Source/GB_AxB_flopcount.c:85-316 (MATLAB statement Test/flopcount.m:12-54)
 * Bflops has B->nvec+1 entries and is returned cumulative. */
int64_t oracle_flopcount (const omat *M, const omat *A, const omat *B, int64_t *Bflops)
{
    int64_t total = 0 ;
    for (int64_t kk = 0 ; kk < B->nvec ; kk++)
    {
        int64_t bjflops = 0 ;
        int64_t pb = B->p [kk], pbe = B->p [kk+1] ;
        int64_t im_first = -1, im_last = -1 ;
        bool go = (pbe > pb) ;
        if (go && M != NULL)
        {
            int64_t pm, pme ;
            if (!lookup (M, vecname (B, kk), &pm, &pme)) go = false ;       /* :211-218 */
            else { im_first = M->i [pm] ; im_last = M->i [pme-1] ; }
        }
        for ( ; go && pb < pbe ; pb++)
        {
            int64_t pa, pe ;
            if (!lookup (A, B->i [pb], &pa, &pe)) continue ;
            if (M != NULL && (A->i [pe-1] < im_first || A->i [pa] > im_last)) continue ;   /* :268 */
            bjflops += pe - pa ;
        }
        if (Bflops) Bflops [kk] = total ;
        total += bjflops ;
    }
    if (Bflops) Bflops [B->nvec] = total ;
    return total ;
}

static int cmp_i64 (const void *a, const void *b)
{
    int64_t x = *(const int64_t *) a, y = *(const int64_t *) b ;
    return (x < y) ? -1 : (x > y) ;
}

/* growable output */
typedef struct { int64_t *i ; char *x ; int64_t n, cap ; size_t zs ; } obuf ;
static int obuf_push (obuf *o, int64_t i, const void *x)
{
    if (o->n == o->cap)
    {
        int64_t cap = o->cap ? 2 * o->cap : 1024 ;
        int64_t *ni = realloc (o->i, cap * sizeof (int64_t)) ;
        if (!ni) return 0 ;
        o->i = ni ;
        char *nx = realloc (o->x, cap * o->zs) ;
        if (!nx) return 0 ;
        o->x = nx ; o->cap = cap ;
    }
    o->i [o->n] = i ;
    memcpy (o->x + o->n * o->zs, x, o->zs) ;
    o->n++ ;
    return 1 ;
}

/* Assemble the result from per-source-vector counts: hypersparse rule of GB_AxB_alloc.c:49-50 and
 * the vector bookkeeping of GB_jstartup/jappend/jwrapup (Source/GB.h:4274-4437). */
static int finish (oresult *R, obuf *o, const int64_t *cnt, int64_t nsrc, const omat *src,
    bool C_is_hyper, int64_t cvlen, int64_t cvdim, int ztype)
{
    R->vlen = cvlen ; R->vdim = cvdim ; R->nnz = o->n ; R->is_hyper = C_is_hyper ;
    R->type_code = ztype ; R->i = o->i ; R->x = o->x ; R->h = NULL ;
    int64_t nonempty = 0 ;
    for (int64_t s = 0 ; s < nsrc ; s++) if (cnt [s] > 0) nonempty++ ;
    R->nvec_nonempty = nonempty ;
    if (C_is_hyper)
    {
        R->nvec = nonempty ;
        R->p = malloc ((nonempty + 1) * sizeof (int64_t)) ;
        R->h = malloc ((nonempty + 1) * sizeof (int64_t)) ;
        if (!R->p || !R->h) return 0 ;
        int64_t q = 0, run = 0 ;
        for (int64_t s = 0 ; s < nsrc ; s++)
            if (cnt [s] > 0) { R->h [q] = vecname (src, s) ; R->p [q++] = run ; run += cnt [s] ; }
        R->p [q] = run ;
    }
    else
    {
        R->nvec = cvdim ;
        R->p = calloc (cvdim + 1, sizeof (int64_t)) ;
        if (!R->p) return 0 ;
        for (int64_t s = 0 ; s < nsrc ; s++) R->p [vecname (src, s) + 1] = cnt [s] ;
        for (int64_t j = 0 ; j < cvdim ; j++) R->p [j+1] += R->p [j] ;
    }
    return 1 ;
}

/* ---- saxpy: GB_AxB_Gustavson (Source/GB_AxB_Gustavson.c:30-432) with the mask policy of
 * GB_AxB_sequential.c:76-95.  Unmasked: symbolic pattern + sort (Gustavson_symbolic.c:187-233) then
 * numeric from the identity in ascending k (Gustavson_nomask.c:91-158).  Masked: valued mask
 * scattered, first product copied, later ones added, gather in mask order (Gustavson_mask.c:153-271,
 * Generator/GB_AxB.c:73-90). */
static int saxpy (oresult *R, const omat *M, int mask_comp, const omat *A, const omat *B, osemiring s)
{
    const size_t xs = tsize [s.xy], zs = tsize [s.z] ;
    const int64_t cvlen = A->vlen, cvdim = B->vdim ;
    if (M != NULL && mask_comp) M = NULL ;
    if (M != NULL)
    {
        int64_t mnz = M->p [M->nvec] ;
        if (oracle_flopcount (M, A, B, NULL) <= mnz) M = NULL ;
    }
    R->mask_applied = (M != NULL) ; R->method_used = 1001 ;
    bool C_is_hyper = (cvdim > 1) && (A->is_hyper || B->is_hyper || (M && M->is_hyper)) ;
    int8_t *mark = calloc (cvlen > 0 ? cvlen : 1, 1) ;
    char *work = malloc ((cvlen > 0 ? cvlen : 1) * zs) ;
    int64_t *cnt = calloc (B->nvec > 0 ? B->nvec : 1, sizeof (int64_t)) ;
    int64_t *pat = malloc ((cvlen > 0 ? cvlen : 1) * sizeof (int64_t)) ;
    obuf o = { NULL, NULL, 0, 0, zs } ;
    if (!mark || !work || !cnt || !pat) return 0 ;
    char t [8] ;
    const char *Ax = A->x, *Bx = B->x ;
    for (int64_t kk = 0 ; kk < B->nvec ; kk++)
    {
        int64_t pb0 = B->p [kk], pb1 = B->p [kk+1] ;
        if (pb1 == pb0) continue ;
        if (M == NULL)
        {
            int64_t n = 0 ;
            for (int64_t pb = pb0 ; pb < pb1 ; pb++)
            {
                int64_t pa, pe ;
                if (!lookup (A, B->i [pb], &pa, &pe)) continue ;
                for (int64_t p = pa ; p < pe ; p++)
                {
                    int64_t i = A->i [p] ;
                    if (!mark [i]) { mark [i] = 1 ; pat [n++] = i ; }
                }
            }
            qsort (pat, n, sizeof (int64_t), cmp_i64) ;
            for (int64_t q = 0 ; q < n ; q++) IDENT [s.z] (s.add, work + pat [q] * zs) ;
            for (int64_t pb = pb0 ; pb < pb1 ; pb++)
            {
                int64_t pa, pe ;
                if (!lookup (A, B->i [pb], &pa, &pe)) continue ;
                for (int64_t p = pa ; p < pe ; p++)
                {
                    MULT [s.xy] (s.mult, Ax + p * xs, Bx + pb * xs, t, s.flip) ;
                    ADD [s.z] (s.add, work + A->i [p] * zs, t) ;
                }
            }
            for (int64_t q = 0 ; q < n ; q++)
            {
                if (!obuf_push (&o, pat [q], work + pat [q] * zs)) return 0 ;
                mark [pat [q]] = 0 ;
            }
            cnt [kk] = n ;
        }
        else
        {
            int64_t pm, pme ;
            if (!lookup (M, vecname (B, kk), &pm, &pme)) continue ;
            for (int64_t p = pm ; p < pme ; p++) if (mask_true (M, p)) mark [M->i [p]] = 1 ;
            for (int64_t pb = pb0 ; pb < pb1 ; pb++)
            {
                int64_t pa, pe ;
                if (!lookup (A, B->i [pb], &pa, &pe)) continue ;
                for (int64_t p = pa ; p < pe ; p++)
                {
                    int64_t i = A->i [p] ;
                    if (!mark [i]) continue ;
                    MULT [s.xy] (s.mult, Ax + p * xs, Bx + pb * xs, t, s.flip) ;
                    if (mark [i] == 1) { mark [i] = 2 ; memcpy (work + i * zs, t, zs) ; }
                    else ADD [s.z] (s.add, work + i * zs, t) ;
                }
            }
            int64_t n = 0 ;
            for (int64_t p = pm ; p < pme ; p++)
            {
                int64_t i = M->i [p] ;
                if (mark [i] == 2) { if (!obuf_push (&o, i, work + i * zs)) return 0 ; n++ ; }
                mark [i] = 0 ;
            }
            cnt [kk] = n ;
        }
    }
    int ok = finish (R, &o, cnt, B->nvec, B, C_is_hyper, cvlen, cvdim, s.z) ;
    free (mark) ; free (work) ; free (cnt) ; free (pat) ;
    return ok ;
}

/* one dot product C(i,j) = A(:,i)'*B(:,j): Source/Template/GB_AxB_dot_cij.c:47-256.  Sparse cases
 * copy the first product and add the rest in ascending k; dense cases start from the identity
 * (:110,125,141). */
static bool dot_cij (const omat *A, const omat *B, int64_t pa, int64_t pe, int64_t pb, int64_t pbe,
    osemiring s, void *cij)
{
    const size_t xs = tsize [s.xy], zs = tsize [s.z] ;
    const char *Ax = A->x, *Bx = B->x ;
    char t [8] ;
    int64_t ainz = pe - pa, bjnz = pbe - pb ;
    if (ainz == 0 || bjnz == 0) return false ;
    if (A->i [pe-1] < B->i [pb] || B->i [pbe-1] < A->i [pa]) return false ;
    if (bjnz == B->vlen || ainz == A->vlen)
    {
        IDENT [s.z] (s.add, cij) ;
        if (ainz == A->vlen && bjnz == B->vlen)
            for (int64_t k = 0 ; k < A->vlen ; k++)
            { MULT [s.xy] (s.mult, Ax + (pa + k) * xs, Bx + (pb + k) * xs, t, s.flip) ; ADD [s.z] (s.add, cij, t) ; }
        else if (ainz == A->vlen)
            for (int64_t p = pb ; p < pbe ; p++)
            { MULT [s.xy] (s.mult, Ax + (pa + B->i [p]) * xs, Bx + p * xs, t, s.flip) ; ADD [s.z] (s.add, cij, t) ; }
        else
            for (int64_t p = pa ; p < pe ; p++)
            { MULT [s.xy] (s.mult, Ax + p * xs, Bx + (pb + A->i [p]) * xs, t, s.flip) ; ADD [s.z] (s.add, cij, t) ; }
        return true ;
    }
    bool found = false ;
    while (pa < pe && pb < pbe)
    {
        int64_t ia = A->i [pa], ib = B->i [pb] ;
        if (ia < ib) pa++ ;
        else if (ib < ia) pb++ ;
        else
        {
            MULT [s.xy] (s.mult, Ax + pa * xs, Bx + pb * xs, t, s.flip) ;
            if (!found) { memcpy (cij, t, zs) ; found = true ; }
            else ADD [s.z] (s.add, cij, t) ;
            pa++ ; pb++ ;
        }
    }
    return found ;
}

/* ---- GB_AxB_dot: Source/GB_AxB_dot.c:39-317 with Template/GB_AxB_dot_mask.c:33-159 (C<M>),
 * dot_compmask.c:20-126 (C<!M>) and dot_nomask.c:20-83 (C) */
static int dot (oresult *R, const omat *M, int mask_comp, const omat *A, const omat *B, osemiring s)
{
    const size_t zs = tsize [s.z] ;
    const int64_t cvlen = A->vdim, cvdim = B->vdim ;
    R->mask_applied = (M != NULL) ; R->method_used = 1003 ;
    bool C_is_hyper = (cvdim > 1) && (A->is_hyper || B->is_hyper || (M && !mask_comp && M->is_hyper)) ;
    int64_t *cnt = calloc (B->nvec > 0 ? B->nvec : 1, sizeof (int64_t)) ;
    obuf o = { NULL, NULL, 0, 0, zs } ;
    if (!cnt) return 0 ;
    char cij [8] ;
    for (int64_t kk = 0 ; kk < B->nvec ; kk++)
    {
        int64_t pb = B->p [kk], pbe = B->p [kk+1] ;
        if (pbe == pb) continue ;
        int64_t j = vecname (B, kk), pm = 0, pme = 0 ;
        if (M != NULL) lookup (M, j, &pm, &pme) ;
        int64_t n = 0 ;
        if (M != NULL && !mask_comp)
        {
            for (int64_t p = pm ; p < pme ; p++)
            {
                if (!mask_true (M, p)) continue ;
                int64_t i = M->i [p], pa, pe ;
                if (!lookup (A, i, &pa, &pe)) continue ;
                if (dot_cij (A, B, pa, pe, pb, pbe, s, cij)) { if (!obuf_push (&o, i, cij)) return 0 ; n++ ; }
            }
        }
        else
        {
            for (int64_t ka = 0 ; ka < A->nvec ; ka++)
            {
                int64_t i = vecname (A, ka) ;
                bool mij = false ;
                if (M != NULL)
                {
                    while (pm < pme && M->i [pm] < i) pm++ ;
                    if (pm < pme && M->i [pm] == i) { mij = mask_true (M, pm) ; pm++ ; }
                }
                if (mij) continue ;
                if (dot_cij (A, B, A->p [ka], A->p [ka+1], pb, pbe, s, cij))
                { if (!obuf_push (&o, i, cij)) return 0 ; n++ ; }
            }
        }
        cnt [kk] = n ;
    }
    int ok = finish (R, &o, cnt, B->nvec, B, C_is_hyper, cvlen, cvdim, s.z) ;
    free (cnt) ;
    return ok ;
}

/* ---- typecasting of built-in types: the generic path of the reference casts every entry of A and B
 * to the multiply operator's input type before use (Source/GB_AxB_Gustavson.c:360-404,
 * GB_AxB_dot.c:189-305, via GB_cast_factory); the rule is GB_CAST, Source/GB.h:2925-2947: a float or
 * double NaN becomes integer 0, +Inf / -Inf the largest / smallest integer, anything else is the C
 * cast; to bool: x != 0.  Casting is a pure function of the entry, so casting the arrays up front
 * gives the same T. */
static void *cast_array (const void *src, int from, int to, int64_t n)
{
    void *dst = malloc ((size_t) (n > 0 ? n : 1) * tsize [to]) ;
    if (dst == NULL) return NULL ;
    for (int64_t k = 0 ; k < n ; k++)
    {
        int64_t sv = 0 ; uint64_t uv = 0 ; double dv = 0 ; float fv = 0 ;
        int kind ;                          /* 0 signed, 1 unsigned, 2 float, 3 double */
        switch (from)
        {
            case T_BOOL   : uv = ((const uint8_t  *) src) [k] ? 1 : 0 ; kind = 1 ; break ;
            case T_INT8   : sv = ((const int8_t   *) src) [k] ; kind = 0 ; break ;
            case T_UINT8  : uv = ((const uint8_t  *) src) [k] ; kind = 1 ; break ;
            case T_INT16  : sv = ((const int16_t  *) src) [k] ; kind = 0 ; break ;
            case T_UINT16 : uv = ((const uint16_t *) src) [k] ; kind = 1 ; break ;
            case T_INT32  : sv = ((const int32_t  *) src) [k] ; kind = 0 ; break ;
            case T_UINT32 : uv = ((const uint32_t *) src) [k] ; kind = 1 ; break ;
            case T_INT64  : sv = ((const int64_t  *) src) [k] ; kind = 0 ; break ;
            case T_UINT64 : uv = ((const uint64_t *) src) [k] ; kind = 1 ; break ;
            case T_FP32   : fv = ((const float    *) src) [k] ; dv = fv ; kind = 2 ; break ;
            default       : dv = ((const double   *) src) [k] ; kind = 3 ; break ;
        }
#define TO_INT(T, LO, HI)                                                                       \
        {                                                                                       \
            T r ;                                                                               \
            if (kind == 0) r = (T) sv ;                                                         \
            else if (kind == 1) r = (T) uv ;                                                    \
            else if (isnan (dv)) r = 0 ;                                                        \
            else if (isinf (dv)) r = (dv > 0) ? HI : LO ;                                       \
            else r = (kind == 2) ? (T) fv : (T) dv ;                                            \
            ((T *) dst) [k] = r ;                                                               \
        }
        switch (to)
        {
            case T_BOOL   :
                ((uint8_t *) dst) [k] = (kind == 0) ? (sv != 0) : ((kind == 1) ? (uv != 0) : (dv != 0)) ;
                break ;
            case T_INT8   : TO_INT (int8_t,   INT8_MIN,  INT8_MAX)   break ;
            case T_UINT8  : TO_INT (uint8_t,  0,         UINT8_MAX)  break ;
            case T_INT16  : TO_INT (int16_t,  INT16_MIN, INT16_MAX)  break ;
            case T_UINT16 : TO_INT (uint16_t, 0,         UINT16_MAX) break ;
            case T_INT32  : TO_INT (int32_t,  INT32_MIN, INT32_MAX)  break ;
            case T_UINT32 : TO_INT (uint32_t, 0,         UINT32_MAX) break ;
            case T_INT64  : TO_INT (int64_t,  INT64_MIN, INT64_MAX)  break ;
            case T_UINT64 : TO_INT (uint64_t, 0,         UINT64_MAX) break ;
            case T_FP32   :
                ((float *) dst) [k] = (kind == 0) ? (float) sv : ((kind == 1) ? (float) uv :
                    ((kind == 2) ? fv : (float) dv)) ;
                break ;
            default       :
                ((double *) dst) [k] = (kind == 0) ? (double) sv : ((kind == 1) ? (double) uv : dv) ;
                break ;
        }
#undef TO_INT
    }
    return dst ;
}

/* ---- entry point: the contract of GB_AxB_parallel (Source/GB.h:1522-1537) on plain arrays.
 * Returns 0 on success, 1 out of memory, 2 semiring not built in. */
int oracle_AxB (oresult *R, const omat *M, int mask_comp, const omat *A, const omat *B,
    int add, int mult, int xy, int z, int flip, int do_adotb)
{
    osemiring s = { add, mult, xy, z, flip } ;
    canonical (&s) ;
    if (A->type_code < T_BOOL || A->type_code > T_FP64 || B->type_code < T_BOOL
        || B->type_code > T_FP64) return 2 ;
    omat Ac = *A, Bc = *B ;
    void *ax = NULL, *bx = NULL ;
    if (A->type_code != s.xy)
    {
        ax = cast_array (A->x, A->type_code, s.xy, A->p [A->nvec]) ;
        if (ax == NULL) return 1 ;
        Ac.x = ax ; Ac.type_code = s.xy ;
    }
    if (B->type_code != s.xy)
    {
        bx = cast_array (B->x, B->type_code, s.xy, B->p [B->nvec]) ;
        if (bx == NULL) { free (ax) ; return 1 ; }
        Bc.x = bx ; Bc.type_code = s.xy ;
    }
    memset (R, 0, sizeof (*R)) ;
    int ok = do_adotb ? dot (R, M, mask_comp, &Ac, &Bc, s) : saxpy (R, M, mask_comp, &Ac, &Bc, s) ;
    free (ax) ; free (bx) ;
    return ok ? 0 : 1 ;
}

/* ---- GB_transpose, the general out-of-place case without an operator (Source/GB_transpose.c:470-985):
 * C = (ctype) A'.  The entries are those of the bucket method (Source/GB_transpose_bucket.c: count the
 * entries per index, cumulative sum, then walk A's vectors in order so that every vector of C comes out
 * ascending); the form of C is the one the reference ends with: the method chosen by the memory estimate
 * of GB_transpose.c:497-606 (quicksort -> GB_builder's hypersparse T, bucket -> standard) followed by
 * GB_to_hyper_conform (Source/GB_to_hyper_conform.c:38-58, tests in GB_to_hyper_test.c and
 * GB_to_nonhyper_test.c, in single precision as there).  Returns 0 on success, 1 out of memory. */
int oracle_transpose (oresult *R, const omat *A, int ctype, double hyper_ratio)
{
    memset (R, 0, sizeof (*R)) ;
    const int64_t anz = A->p [A->nvec], avlen = A->vlen, avdim = A->vdim ;
    const size_t csize = tsize [ctype] ;
    /* the form C starts out in */
    bool c_hyper = true ;
    if (!A->is_hyper)
    {
        double q = 0, qmax = 0 ;
        q += ((double) anz * 8.0) / 1e9 ; q += ((double) anz * 8.0) / 1e9 ; qmax = q ;
        q += ((double) anz * 8.0) / 1e9 ; if (q > qmax) qmax = q ;
        q -= ((double) anz * 8.0) / 1e9 ; q += ((double) anz * (double) csize) / 1e9 ; if (q > qmax) qmax = q ;
        double b = ((double) avlen * 8.0) / 1e9 + ((double) anz * 8.0) / 1e9
            + ((double) anz * (double) csize) / 1e9 + ((double) avlen * 8.0) / 1e9 ;
        c_hyper = (qmax < b) ;
    }
    int64_t *cnt = calloc ((size_t) avlen + 1, sizeof (int64_t)) ;
    int64_t *w = malloc (((size_t) avlen + 1) * sizeof (int64_t)) ;
    int64_t *Ci = malloc ((size_t) (anz > 0 ? anz : 1) * sizeof (int64_t)) ;
    int64_t *src = malloc ((size_t) (anz > 0 ? anz : 1) * sizeof (int64_t)) ;
    if (!cnt || !w || !Ci || !src) { free (cnt) ; free (w) ; free (Ci) ; free (src) ; return 1 ; }
    for (int64_t p = 0 ; p < anz ; p++) cnt [A->i [p]]++ ;
    int64_t run = 0, nonempty = 0 ;
    for (int64_t i = 0 ; i < avlen ; i++) { w [i] = run ; run += cnt [i] ; if (cnt [i] > 0) nonempty++ ; }
    for (int64_t k = 0 ; k < A->nvec ; k++)
    {
        const int64_t j = vecname (A, k) ;
        for (int64_t p = A->p [k] ; p < A->p [k+1] ; p++)
        {
            const int64_t q = w [A->i [p]]++ ;
            Ci [q] = j ; src [q] = p ;
        }
    }
    /* values, cast entry by entry (GB_cast_array) */
    void *cast = (ctype != A->type_code) ? cast_array (A->x, A->type_code, ctype, anz) : NULL ;
    char *Cx = malloc ((size_t) (anz > 0 ? anz : 1) * csize) ;
    if (Cx == NULL || (ctype != A->type_code && cast == NULL))
    { free (cnt) ; free (w) ; free (Ci) ; free (src) ; free (cast) ; free (Cx) ; return 1 ; }
    const char *from = (cast != NULL) ? (const char *) cast : (const char *) A->x ;
    for (int64_t q = 0 ; q < anz ; q++) memcpy (Cx + q * csize, from + src [q] * csize, csize) ;
    free (cast) ; free (src) ; free (w) ;
    /* GB_to_hyper_conform */
    {
        const float n = (float) avlen, r = (float) hyper_ratio ;
        const float k = (float) nonempty ;
        if (!c_hyper) { if (n > 1 && k <= n * r) c_hyper = true ; }
        else if (n <= 1 || k > n * r * 2) c_hyper = false ;
    }
    R->vlen = avdim ; R->vdim = avlen ; R->nnz = anz ; R->is_hyper = c_hyper ; R->type_code = ctype ;
    R->i = Ci ; R->x = Cx ; R->nvec_nonempty = nonempty ;
    if (c_hyper)
    {
        R->nvec = nonempty ;
        R->p = malloc (((size_t) nonempty + 1) * sizeof (int64_t)) ;
        R->h = malloc (((size_t) nonempty + 1) * sizeof (int64_t)) ;
        if (!R->p || !R->h) { free (cnt) ; return 1 ; }
        int64_t q = 0 ; run = 0 ;
        for (int64_t i = 0 ; i < avlen ; i++)
            if (cnt [i] > 0) { R->h [q] = i ; R->p [q++] = run ; run += cnt [i] ; }
        R->p [q] = run ;
    }
    else
    {
        R->nvec = avlen ;
        R->p = malloc (((size_t) avlen + 1) * sizeof (int64_t)) ;
        if (!R->p) { free (cnt) ; return 1 ; }
        run = 0 ;
        for (int64_t i = 0 ; i < avlen ; i++) { R->p [i] = run ; run += cnt [i] ; }
        R->p [avlen] = run ;
    }
    free (cnt) ;
    return 0 ;
}

/* ---- GB_accum_mask: C<M> = accum (C,T) (Source/GB_accum_mask.c:130-328), restated vector by vector as
 * the three-way merge of GB_add.c (Z = accum (C,T): both -> the operator on C cast to x and T cast to y, the
 * result cast to C's type; only C -> C; only T -> T cast to C's type; no accumulator: Z = T cast to C's type)
 * followed by GB_mask.c (GB_spec_mask.m:60-90: where the mask admits, the entry of Z or none; elsewhere the
 * entry of C, or none under C_replace; a mask entry admits iff its value cast to bool is true, negated for
 * a complemented mask; no mask admits everywhere).  R is hypersparse iff r_hyper (GB_mask.c:315: C and Z
 * both hypersparse).  accum_op == 0: no accumulator.  Returns 0 on success, 1 out of memory. */
int oracle_accum_mask (oresult *R, const omat *C, const omat *T, const omat *M, int mask_comp, int replace,
    int accum_op, int accum_xy, int r_hyper)
{
    memset (R, 0, sizeof (*R)) ;
    const int ctype = C->type_code ;
    const size_t cs = tsize [ctype] ;
    const int64_t cnz = C->p [C->nvec], tnz = T->p [T->nvec] ;
    int op = accum_op ;
    if (op != 0 && accum_xy == T_BOOL) op = boolean_rename (op) ;
    const int ztype = (op == 0) ? ctype : ((accum_xy == T_BOOL || op >= OP_EQ) ? T_BOOL : accum_xy) ;
    void *Tc = cast_array (T->x, T->type_code, ctype, tnz) ;
    void *Cx = (op != 0) ? cast_array (C->x, ctype, accum_xy, cnz) : NULL ;
    void *Ty = (op != 0) ? cast_array (T->x, T->type_code, accum_xy, tnz) : NULL ;
    int64_t *cnt = calloc ((size_t) C->vdim + 1, sizeof (int64_t)) ;
    obuf o = { NULL, NULL, 0, 0, cs } ;
    if (!Tc || !cnt || (op != 0 && (!Cx || !Ty))) { free (Tc) ; free (Cx) ; free (Ty) ; free (cnt) ; return 1 ; }
    const size_t xs = (op != 0) ? tsize [accum_xy] : 0 ;
    int ok = 1 ;
    for (int64_t j = 0 ; j < C->vdim && ok ; j++)
    {
        int64_t pc, pce, pt, pte, pm = 0, pme = 0 ;
        lookup (C, j, &pc, &pce) ;
        lookup (T, j, &pt, &pte) ;
        if (M != NULL) lookup (M, j, &pm, &pme) ;
        while ((pc < pce || pt < pte) && ok)
        {
            const int64_t ic = (pc < pce) ? C->i [pc] : INT64_MAX, it = (pt < pte) ? T->i [pt] : INT64_MAX ;
            const int64_t i = (ic < it) ? ic : it ;
            const bool cex = (ic == i), tex = (it == i) ;
            bool m = true ;
            if (M != NULL)
            {
                while (pm < pme && M->i [pm] < i) pm++ ;
                m = (pm < pme && M->i [pm] == i && mask_true (M, pm)) ;
            }
            if (mask_comp) m = !m ;
            char z [16], zc [16] ;
            const void *val = NULL ;
            if (m)
            {
                if (op == 0) { if (tex) val = (char *) Tc + pt * cs ; }
                else if (cex && tex)
                {
                    MULT [accum_xy] (op, (char *) Cx + pc * xs, (char *) Ty + pt * xs, z, 0) ;
                    void *one = cast_array (z, ztype, ctype, 1) ;
                    if (one == NULL) { ok = 0 ; break ; }
                    memcpy (zc, one, cs) ;
                    free (one) ;
                    val = zc ;
                }
                else if (cex) val = (const char *) C->x + pc * cs ;
                else val = (char *) Tc + pt * cs ;
            }
            else if (!replace && cex) val = (const char *) C->x + pc * cs ;
            if (val != NULL)
            {
                if (!obuf_push (&o, i, val)) { ok = 0 ; break ; }
                cnt [j]++ ;
            }
            if (cex) pc++ ;
            if (tex) pt++ ;
        }
    }
    free (Tc) ; free (Cx) ; free (Ty) ;
    if (ok)
    {
        omat all = { C->vlen, C->vdim, C->vdim, NULL, NULL, NULL, NULL, ctype, 0 } ;
        ok = finish (R, &o, cnt, C->vdim, &all, r_hyper != 0, C->vlen, C->vdim, ctype) ;
    }
    free (cnt) ;
    if (!ok) { free (o.i) ; free (o.x) ; return 1 ; }
    return 0 ;
}

void oracle_free (oresult *R)
{
    free (R->p) ; free (R->h) ; free (R->i) ; free (R->x) ;
    memset (R, 0, sizeof (*R)) ;
}
