// inst_int8.cu -- semiring kernels for operands of type int8_t (see kernels.cuh)
#include "kernels.cuh"
GB200_INSTANTIATE_TYPE (int8, int8_t)
