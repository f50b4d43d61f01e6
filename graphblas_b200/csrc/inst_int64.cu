// inst_int64.cu -- semiring kernels for operands of type int64_t (see kernels.cuh)
#include "kernels.cuh"
GB200_INSTANTIATE_TYPE (int64, int64_t)
