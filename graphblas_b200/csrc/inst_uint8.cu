// inst_uint8.cu -- semiring kernels for operands of type uint8_t (see kernels.cuh)
#include "kernels.cuh"
GB200_INSTANTIATE_TYPE (uint8, uint8_t)
