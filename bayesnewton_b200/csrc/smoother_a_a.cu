// RTS smoother instantiations, array-entry group A (see common.cuh).
#include "smoother_impl.cuh"
namespace bn {
int rts_group_a_a(const RtsCall& c) {
    BN_GROUP_A_A(BN_RTS_ARR_CASE)
    return kNotHandled;
}
}  // namespace bn
