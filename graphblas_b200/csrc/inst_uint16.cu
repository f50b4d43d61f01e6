// inst_uint16.cu -- semiring kernels for operands of type uint16_t (see kernels.cuh)
#include "kernels.cuh"
GB200_INSTANTIATE_TYPE (uint16, uint16_t)
