// RTS smoother instantiations, array-entry group C (see common.cuh).
#include "smoother_impl.cuh"
namespace bn {
int rts_group_a_c(const RtsCall& c) {
    BN_GROUP_A_C(BN_RTS_ARR_CASE)
    return kNotHandled;
}
}  // namespace bn
