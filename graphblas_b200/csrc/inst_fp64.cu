// inst_fp64.cu -- semiring kernels for operands of type double (see kernels.cuh)
#include "kernels.cuh"
GB200_INSTANTIATE_TYPE (fp64, double)
