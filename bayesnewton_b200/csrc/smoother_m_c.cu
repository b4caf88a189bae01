// RTS smoother instantiations, stationary-kernel group C (see common.cuh).
#include "smoother_impl.cuh"
namespace bn {
int rts_group_m_c(const RtsCall& c) {
    BN_GROUP_M_C(BN_RTS_SPEC_CASE)
    return kNotHandled;
}
}  // namespace bn
