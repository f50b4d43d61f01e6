// inst_int32.cu -- semiring kernels for operands of type int32_t (see kernels.cuh)
#include "kernels.cuh"
GB200_INSTANTIATE_TYPE (int32, int32_t)
