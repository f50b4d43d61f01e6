// inst_uint32.cu -- semiring kernels for operands of type uint32_t (see kernels.cuh)
#include "kernels.cuh"
GB200_INSTANTIATE_TYPE (uint32, uint32_t)
