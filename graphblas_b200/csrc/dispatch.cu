// dispatch.cu -- run-time (xy type) -> compiled launcher table.  Each inst_<type>.cu translation
// unit instantiates the semiring-templated kernels for one operand type, so they build in parallel.
#include "engine.cuh"
#include "kernels.cuh"

namespace gb200 {

#define GB200_DECL(NAME) bool launch_##NAME (int family, int z_code, int add, int mult, \
    const void *args, LaunchCfg cfg) ;
GB200_DECL (bool)   GB200_DECL (int8)   GB200_DECL (uint8)  GB200_DECL (int16)
GB200_DECL (uint16) GB200_DECL (int32)  GB200_DECL (uint32) GB200_DECL (int64)
GB200_DECL (uint64) GB200_DECL (fp32)   GB200_DECL (fp64)
#undef GB200_DECL

static void next_kev (Ctx &c)
{
    if (c.kev_used + 2 > (int) c.kev.size ())
    {
        cudaEvent_t e0, e1 ;
        cudaEventCreate (&e0) ; cudaEventCreate (&e1) ;
        c.kev.push_back (e0) ; c.kev.push_back (e1) ;
    }
}

gb200_status group_begin ()
{
    Ctx &c = ctx () ;
    next_kev (c) ;
    GB200_CUDA (cudaEventRecord (c.kev [c.kev_used], c.stream)) ;
    GB200_CUDA (cudaEventRecord (c.fork_ev, c.stream)) ;
    for (int k = 0 ; k < Ctx::NSIDE ; k++) GB200_CUDA (cudaStreamWaitEvent (c.side [k], c.fork_ev, 0)) ;
    c.group_stream = c.side [0] ;
    return GB200_SUCCESS ;
}

void group_use (int k) { Ctx &c = ctx () ; c.group_stream = c.side [k % Ctx::NSIDE] ; }

gb200_status group_end ()
{
    Ctx &c = ctx () ;
    c.group_stream = nullptr ;
    for (int k = 0 ; k < Ctx::NSIDE ; k++)
    {
        GB200_CUDA (cudaEventRecord (c.side_done [k], c.side [k])) ;
        GB200_CUDA (cudaStreamWaitEvent (c.stream, c.side_done [k], 0)) ;
    }
    GB200_CUDA (cudaEventRecord (c.kev [c.kev_used + 1], c.stream)) ;
    c.kev_used += 2 ;
    return GB200_SUCCESS ;
}

bool launch_typed (int xy_code, int family, int z_code, int add, int mult, const void *args,
    int grid, int block)
{
    LaunchCfg cfg ;
    Ctx &c = ctx () ;
    const bool grouped = (c.group_stream != nullptr) ;
    cfg.grid = grid ; cfg.block = block ; cfg.stream = grouped ? c.group_stream : c.stream ;
    if (!grouped)
    {
        next_kev (c) ;
        cudaEventRecord (c.kev [c.kev_used], c.stream) ;
    }
    bool ok = false ;
    switch (xy_code)
    {
        case GB200_BOOL   : ok = launch_bool   (family, z_code, add, mult, args, cfg) ; break ;
        case GB200_INT8   : ok = launch_int8   (family, z_code, add, mult, args, cfg) ; break ;
        case GB200_UINT8  : ok = launch_uint8  (family, z_code, add, mult, args, cfg) ; break ;
        case GB200_INT16  : ok = launch_int16  (family, z_code, add, mult, args, cfg) ; break ;
        case GB200_UINT16 : ok = launch_uint16 (family, z_code, add, mult, args, cfg) ; break ;
        case GB200_INT32  : ok = launch_int32  (family, z_code, add, mult, args, cfg) ; break ;
        case GB200_UINT32 : ok = launch_uint32 (family, z_code, add, mult, args, cfg) ; break ;
        case GB200_INT64  : ok = launch_int64  (family, z_code, add, mult, args, cfg) ; break ;
        case GB200_UINT64 : ok = launch_uint64 (family, z_code, add, mult, args, cfg) ; break ;
        case GB200_FP32   : ok = launch_fp32   (family, z_code, add, mult, args, cfg) ; break ;
        case GB200_FP64   : ok = launch_fp64   (family, z_code, add, mult, args, cfg) ; break ;
        default : break ;
    }
    if (!grouped)
    {
        cudaEventRecord (c.kev [c.kev_used + 1], c.stream) ;
        c.kev_used += 2 ;
    }
    if (ok) count_launch () ;
    return ok ;
}

} // namespace gb200
