// inst_uint64.cu -- semiring kernels for operands of type uint64_t (see kernels.cuh)
#include "kernels.cuh"
GB200_INSTANTIATE_TYPE (uint64, uint64_t)
