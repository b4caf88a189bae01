// RTS smoother instantiations, array-entry group B (see common.cuh).
#include "smoother_impl.cuh"
namespace bn {
int rts_group_a_b(const RtsCall& c) {
    BN_GROUP_A_B(BN_RTS_ARR_CASE)
    return kNotHandled;
}
}  // namespace bn
