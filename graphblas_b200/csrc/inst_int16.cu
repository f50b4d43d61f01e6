// inst_int16.cu -- semiring kernels for operands of type int16_t (see kernels.cuh)
#include "kernels.cuh"
GB200_INSTANTIATE_TYPE (int16, int16_t)
