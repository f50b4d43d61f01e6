// inst_fp32.cu -- semiring kernels for operands of type float (see kernels.cuh)
#include "kernels.cuh"
GB200_INSTANTIATE_TYPE (fp32, float)
