// fp32 entry points of the fused iteration (bn_iter_*_f32): iter.cu compiled with a float scalar type
#define BN_REAL32 1
#define BN_NS bn32
#include "iter.cu"
