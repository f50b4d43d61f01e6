// inst_bool.cu -- semiring kernels for operands of type bool (see kernels.cuh)
#include "kernels.cuh"
GB200_INSTANTIATE_TYPE (bool, bool)
