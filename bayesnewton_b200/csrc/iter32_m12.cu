// fp32 build of the fused iteration (real.cuh): the same kernels with a float scalar type, in namespace bn32
#define BN_REAL32 1
#define BN_NS bn32
#include "iter_m12.cu"
