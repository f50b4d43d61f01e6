// semiring.cuh -- the built-in semiring space of SuiteSparse:GraphBLAS v2.3.3 as device functors.
//
// The reference instantiates 960 (add, mult, type) workers by macro expansion
// (Source/Generator/GB_AxB.c:30-222, tables Source/axb.m:19-45, Source/axb_template.m:25-86,
// Source/axb_compare_template.m:4-66).  Here the same space is two small template families:
//   gb_mult<T,Z,OP>(x,y,flip)  -- the 23 multiply operators (OP < 0: uniform run-time switch)
//   Monoid<Z,ADD>              -- the 8 monoids: identity, terminal, combine, atomic_combine
// Arithmetic follows the C semantics of the generated code: narrow integers are computed in
// int and truncated, signed overflow wraps, integer division by zero follows GB_IDIV
// (Source/GB.h:2782-2870), float min/max are fmin/fmax (omit-NaN), and this file must be
// compiled with -fmad=false because the reference's x86-64 build emits no FMA.
#pragma once
#include <cstdint>
#include <cfloat>
#include <climits>
#include <cmath>
#include <type_traits>
#include <cuda_runtime.h>
#include "../../include/gb_b200.h"

namespace gb200 {

// ---------------------------------------------------------------------------------------------
// type code <-> C type
// ---------------------------------------------------------------------------------------------
template <int CODE> struct TypeOf ;
template <> struct TypeOf<GB200_BOOL>   { using type = bool ; } ;
template <> struct TypeOf<GB200_INT8>   { using type = int8_t ; } ;
template <> struct TypeOf<GB200_UINT8>  { using type = uint8_t ; } ;
template <> struct TypeOf<GB200_INT16>  { using type = int16_t ; } ;
template <> struct TypeOf<GB200_UINT16> { using type = uint16_t ; } ;
template <> struct TypeOf<GB200_INT32>  { using type = int32_t ; } ;
template <> struct TypeOf<GB200_UINT32> { using type = uint32_t ; } ;
template <> struct TypeOf<GB200_INT64>  { using type = int64_t ; } ;
template <> struct TypeOf<GB200_UINT64> { using type = uint64_t ; } ;
template <> struct TypeOf<GB200_FP32>   { using type = float ; } ;
template <> struct TypeOf<GB200_FP64>   { using type = double ; } ;

__host__ __device__ inline int type_size (int code)
{
    switch (code)
    {
        case GB200_BOOL: case GB200_INT8: case GB200_UINT8: return 1 ;
        case GB200_INT16: case GB200_UINT16: return 2 ;
        case GB200_INT32: case GB200_UINT32: case GB200_FP32: return 4 ;
        default: return 8 ;
    }
}

// The accumulator type: values narrower than 32 bits are accumulated in a 32-bit word (the
// hardware has no 8/16-bit atomics) and truncated once at the end.  For PLUS and TIMES this is
// exact because truncation commutes with wrap-around +,*; for MIN/MAX/LOR/LAND/LXOR/EQ the
// widened value is the same number.
template <class Z> struct AccOf { using type = Z ; } ;
template <> struct AccOf<bool>     { using type = uint32_t ; } ;
template <> struct AccOf<int8_t>   { using type = int32_t ; } ;
template <> struct AccOf<uint8_t>  { using type = uint32_t ; } ;
template <> struct AccOf<int16_t>  { using type = int32_t ; } ;
template <> struct AccOf<uint16_t> { using type = uint32_t ; } ;

template <class Z> __host__ __device__ inline Z acc_to_z (typename AccOf<Z>::type a)
{
    if constexpr (std::is_same<Z, bool>::value) return (a != 0) ;
    else return (Z) a ;
}

// ---------------------------------------------------------------------------------------------
// numeric limits usable on the device (GB_PLUS_INF / GB_MINUS_INF, Source/GB.h:2593-2650)
// ---------------------------------------------------------------------------------------------
template <class T> struct Lim ;
#define GB200_LIM(T, LO, HI) template <> struct Lim<T> { \
    __host__ __device__ static constexpr T lo () { return LO ; } \
    __host__ __device__ static constexpr T hi () { return HI ; } } ;
GB200_LIM (int8_t,   INT8_MIN,  INT8_MAX)
GB200_LIM (uint8_t,  0,         UINT8_MAX)
GB200_LIM (int16_t,  INT16_MIN, INT16_MAX)
GB200_LIM (uint16_t, 0,         UINT16_MAX)
GB200_LIM (int32_t,  INT32_MIN, INT32_MAX)
GB200_LIM (uint32_t, 0,         UINT32_MAX)
GB200_LIM (int64_t,  INT64_MIN, INT64_MAX)
GB200_LIM (uint64_t, 0,         UINT64_MAX)
#undef GB200_LIM
template <> struct Lim<float>  {
    __host__ __device__ static float lo () { return -INFINITY ; }
    __host__ __device__ static float hi () { return  INFINITY ; } } ;
template <> struct Lim<double> {
    __host__ __device__ static double lo () { return -(double) INFINITY ; }
    __host__ __device__ static double hi () { return  (double) INFINITY ; } } ;

// ---------------------------------------------------------------------------------------------
// integer helpers with the reference's semantics
// ---------------------------------------------------------------------------------------------
template <class T> __host__ __device__ inline T wrap_add (T x, T y)
{
    if constexpr (std::is_floating_point<T>::value) return x + y ;
    else { using U = typename std::make_unsigned<T>::type ; return (T) ((U) x + (U) y) ; }
}
template <class T> __host__ __device__ inline T wrap_sub (T x, T y)
{
    if constexpr (std::is_floating_point<T>::value) return x - y ;
    else { using U = typename std::make_unsigned<T>::type ; return (T) ((U) x - (U) y) ; }
}
template <class T> __host__ __device__ inline T wrap_mul (T x, T y)
{
    if constexpr (std::is_floating_point<T>::value) return x * y ;
    else if constexpr (sizeof (T) < 4) return (T) ((int) x * (int) y) ;
    else { using U = typename std::make_unsigned<T>::type ; return (T) ((U) x * (U) y) ; }
}
// GB_IDIV, Source/GB.h:2782-2870: x/0 = 0 if x == 0, type max if x > 0, type min if x < 0;
// signed x/(-1) = -x (wraps for the most negative value)
template <class T> __host__ __device__ inline T gb_div (T x, T y)
{
    if constexpr (std::is_floating_point<T>::value) return x / y ;
    else if constexpr (std::is_signed<T>::value)
    {
        if (y == (T) (-1)) return wrap_sub<T> ((T) 0, x) ;
        if (y == 0) return (x == 0) ? (T) 0 : ((x < 0) ? Lim<T>::lo () : Lim<T>::hi ()) ;
        return (T) (x / y) ;
    }
    else
    {
        if (y == 0) return (x == 0) ? (T) 0 : Lim<T>::hi () ;
        return (T) (x / y) ;
    }
}
template <class T> __host__ __device__ inline T gb_min (T x, T y)
{
    if constexpr (std::is_same<T, float>::value) return fminf (x, y) ;
    else if constexpr (std::is_same<T, double>::value) return fmin (x, y) ;
    else return (x < y) ? x : y ;                       // GB_IMIN, Source/GB.h:3377
}
template <class T> __host__ __device__ inline T gb_max (T x, T y)
{
    if constexpr (std::is_same<T, float>::value) return fmaxf (x, y) ;
    else if constexpr (std::is_same<T, double>::value) return fmax (x, y) ;
    else return (x > y) ? x : y ;                       // GB_IMAX, Source/GB.h:3378
}

// ---------------------------------------------------------------------------------------------
// multiply operators.  z = mult (x,y); `flip` only matters for MINUS and DIV (Source/axb.m:25,27)
// because gb200_semiring_canonical has already folded it into every other opcode.
// ---------------------------------------------------------------------------------------------
template <class T, class Z> __host__ __device__ inline Z mult_one (int op, T x, T y, bool flip)
{
    if constexpr (std::is_same<T, bool>::value)
    {
        // after the boolean renames only these survive (Source/GB_boolean_rename.c:36-90)
        bool r ;
        switch (op)
        {
            case GB200_FIRST  : r = x ; break ;
            case GB200_SECOND : r = y ; break ;
            case GB200_LOR    : r = x || y ; break ;
            case GB200_LAND   : r = x && y ; break ;
            case GB200_LXOR   : r = (x != y) ; break ;
            case GB200_EQ     : r = (x == y) ; break ;
            case GB200_GT     : r = (x >  y) ; break ;
            case GB200_LT     : r = (x <  y) ; break ;
            case GB200_GE     : r = (x >= y) ; break ;
            default           : r = (x <= y) ; break ;     // GB200_LE
        }
        return (Z) r ;
    }
    else if constexpr (std::is_same<Z, bool>::value)
    {
        switch (op)                                         // TxT -> bool comparators
        {
            case GB200_EQ : return (x == y) ;
            case GB200_NE : return (x != y) ;
            case GB200_GT : return (x >  y) ;
            case GB200_LT : return (x <  y) ;
            case GB200_GE : return (x >= y) ;
            default       : return (x <= y) ;               // GB200_LE
        }
    }
    else
    {
        switch (op)                                         // TxT -> T
        {
            case GB200_FIRST  : return (Z) x ;
            case GB200_SECOND : return (Z) y ;
            case GB200_MIN    : return (Z) gb_min<T> (x, y) ;
            case GB200_MAX    : return (Z) gb_max<T> (x, y) ;
            case GB200_PLUS   : return (Z) wrap_add<T> (x, y) ;
            case GB200_MINUS  : return (Z) (flip ? wrap_sub<T> (y, x) : wrap_sub<T> (x, y)) ;
            case GB200_TIMES  : return (Z) wrap_mul<T> (x, y) ;
            case GB200_DIV    : return (Z) (flip ? gb_div<T> (y, x) : gb_div<T> (x, y)) ;
            case GB200_ISEQ   : return (Z) (x == y) ;
            case GB200_ISNE   : return (Z) (x != y) ;
            case GB200_ISGT   : return (Z) (x >  y) ;
            case GB200_ISLT   : return (Z) (x <  y) ;
            case GB200_ISGE   : return (Z) (x >= y) ;
            case GB200_ISLE   : return (Z) (x <= y) ;
            case GB200_LOR    : return (Z) ((x != 0) || (y != 0)) ;
            case GB200_LAND   : return (Z) ((x != 0) && (y != 0)) ;
            default           : return (Z) ((x != 0) != (y != 0)) ;    // GB200_LXOR
        }
    }
}

// OP >= 0: compile-time operator (the switch folds away); OP < 0: uniform run-time switch
template <class T, class Z, int OP> struct Mult
{
    int op ; bool flip ;
    __host__ __device__ Mult (int op_, bool flip_) : op (op_), flip (flip_) { }
    __host__ __device__ inline Z operator() (T x, T y) const
    {
        if constexpr (OP >= 0) return mult_one<T, Z> (OP, x, y, flip) ;
        else return mult_one<T, Z> (op, x, y, flip) ;
    }
} ;

// ---------------------------------------------------------------------------------------------
// monoids (Source/axb_template.m:25-86): identity, terminal, combine, atomic combine
// ---------------------------------------------------------------------------------------------
template <class A> __device__ inline A atomic_cas (A *addr, A expect, A val)
{
    if constexpr (sizeof (A) == 4)
    {
        unsigned int r = atomicCAS ((unsigned int *) addr, *(unsigned int *) &expect,
            *(unsigned int *) &val) ;
        return *(A *) &r ;
    }
    else
    {
        unsigned long long r = atomicCAS ((unsigned long long *) addr,
            *(unsigned long long *) &expect, *(unsigned long long *) &val) ;
        return *(A *) &r ;
    }
}

template <class Z, int ADD> struct Monoid
{
    using acc_t = typename AccOf<Z>::type ;
    static constexpr int add = ADD ;

    __host__ __device__ static inline acc_t identity ()
    {
        if constexpr (ADD == GB200_MIN) return (acc_t) Lim<Z>::hi () ;
        else if constexpr (ADD == GB200_MAX) return (acc_t) Lim<Z>::lo () ;
        else if constexpr (ADD == GB200_PLUS) return (acc_t) 0 ;
        else if constexpr (ADD == GB200_TIMES) return (acc_t) 1 ;
        else if constexpr (ADD == GB200_LOR) return (acc_t) 0 ;
        else if constexpr (ADD == GB200_LAND) return (acc_t) 1 ;
        else if constexpr (ADD == GB200_LXOR) return (acc_t) 0 ;
        else return (acc_t) 1 ;                                         // EQ
    }
    // terminal value: reaching it allows the dot product to stop (never changes a result)
    __host__ __device__ static constexpr bool has_terminal ()
    {
        return (ADD == GB200_MIN || ADD == GB200_MAX || ADD == GB200_LOR || ADD == GB200_LAND
            || (ADD == GB200_TIMES && !std::is_floating_point<Z>::value)) ;
    }
    __host__ __device__ static inline bool is_terminal (acc_t a)
    {
        // the accumulator may be the widened (untruncated) value, so test the truncated one
        if constexpr (ADD == GB200_MIN) return acc_to_z<Z> (a) == Lim<Z>::lo () ;
        else if constexpr (ADD == GB200_MAX) return acc_to_z<Z> (a) == Lim<Z>::hi () ;
        else if constexpr (ADD == GB200_TIMES && !std::is_floating_point<Z>::value)
            return acc_to_z<Z> (a) == (Z) 0 ;
        else if constexpr (ADD == GB200_LOR) return a != 0 ;
        else if constexpr (ADD == GB200_LAND) return a == 0 ;
        else return false ;
    }
    __host__ __device__ static inline acc_t widen (Z t) { return (acc_t) t ; }

    // w = add (w, t) on plain values
    __host__ __device__ static inline acc_t combine (acc_t w, acc_t t)
    {
        if constexpr (ADD == GB200_MIN) return gb_min<acc_t> (w, t) ;
        else if constexpr (ADD == GB200_MAX) return gb_max<acc_t> (w, t) ;
        else if constexpr (ADD == GB200_PLUS) return wrap_add<acc_t> (w, t) ;
        else if constexpr (ADD == GB200_TIMES) return wrap_mul<acc_t> (w, t) ;
        else if constexpr (ADD == GB200_LOR) return (acc_t) ((w != 0) || (t != 0)) ;
        else if constexpr (ADD == GB200_LAND) return (acc_t) ((w != 0) && (t != 0)) ;
        else if constexpr (ADD == GB200_LXOR) return (acc_t) ((w != 0) != (t != 0)) ;
        else return (acc_t) ((w != 0) == (t != 0)) ;                    // EQ
    }

    // *w = add (*w, t), atomically (shared or global memory).  Integer and boolean monoids are
    // associative and commutative bit-for-bit, so any order gives the reference's result.
    __device__ static inline void atomic_combine (acc_t *w, acc_t t)
    {
        if constexpr (ADD == GB200_PLUS)
        {
            if constexpr (std::is_same<acc_t, int64_t>::value || std::is_same<acc_t, uint64_t>::value)
                atomicAdd ((unsigned long long *) w, (unsigned long long) t) ;
            else if constexpr (std::is_same<acc_t, int32_t>::value)
                atomicAdd ((int *) w, (int) t) ;
            else if constexpr (std::is_same<acc_t, uint32_t>::value)
                atomicAdd ((unsigned int *) w, (unsigned int) t) ;
            else atomicAdd (w, t) ;                                     // float, double
        }
        else if constexpr ((ADD == GB200_MIN || ADD == GB200_MAX)
            && !std::is_floating_point<acc_t>::value)
        {
            if constexpr (std::is_same<acc_t, int64_t>::value)
            {
                if constexpr (ADD == GB200_MIN) atomicMin ((long long *) w, (long long) t) ;
                else atomicMax ((long long *) w, (long long) t) ;
            }
            else if constexpr (std::is_same<acc_t, uint64_t>::value)
            {
                if constexpr (ADD == GB200_MIN) atomicMin ((unsigned long long *) w, (unsigned long long) t) ;
                else atomicMax ((unsigned long long *) w, (unsigned long long) t) ;
            }
            else if constexpr (std::is_same<acc_t, int32_t>::value)
            {
                if constexpr (ADD == GB200_MIN) atomicMin ((int *) w, (int) t) ;
                else atomicMax ((int *) w, (int) t) ;
            }
            else
            {
                if constexpr (ADD == GB200_MIN) atomicMin ((unsigned int *) w, (unsigned int) t) ;
                else atomicMax ((unsigned int *) w, (unsigned int) t) ;
            }
        }
        else if constexpr (ADD == GB200_LOR)  { if (t != 0) atomicOr  ((unsigned int *) w, 1u) ; }
        else if constexpr (ADD == GB200_LAND) { if (t == 0) atomicAnd ((unsigned int *) w, 0u) ; }
        else if constexpr (ADD == GB200_LXOR) { if (t != 0) atomicXor ((unsigned int *) w, 1u) ; }
        else if constexpr (ADD == GB200_EQ)   { if (t == 0) atomicXor ((unsigned int *) w, 1u) ; }
        else
        {
            // TIMES (all types) and floating-point MIN/MAX: compare-and-swap loop
            acc_t old = *((volatile acc_t *) w) ;
            while (true)
            {
                acc_t nw = combine (old, t) ;
                if (memcmp_eq (nw, old)) break ;
                acc_t seen = atomic_cas<acc_t> (w, old, nw) ;
                if (memcmp_eq (seen, old)) break ;
                old = seen ;
            }
        }
    }
    __host__ __device__ static inline bool memcmp_eq (acc_t a, acc_t b)
    {
        if constexpr (sizeof (acc_t) == 4) return *(uint32_t *) &a == *(uint32_t *) &b ;
        else return *(uint64_t *) &a == *(uint64_t *) &b ;
    }
} ;

// A semiring = monoid + multiply; T: type of A and B values, Z: type of the result
template <class T_, class Z_, int ADD, int MULT> struct Semiring
{
    using T = T_ ; using Z = Z_ ;
    using Mon = Monoid<Z_, ADD> ;
    using acc_t = typename Mon::acc_t ;
    Mult<T_, Z_, MULT> mult ;
    __host__ __device__ Semiring (int op, bool flip) : mult (op, flip) { }
    // a is the entry of the first matrix (A), b of the second (B)
    __host__ __device__ inline acc_t product (T a, T b) const { return Mon::widen (mult (a, b)) ; }
} ;

} // namespace gb200
