// RTS smoother instantiations, stationary-kernel group D (see common.cuh).
#include "smoother_impl.cuh"
namespace bn {
int rts_group_m_d(const RtsCall& c) {
    BN_GROUP_M_D(BN_RTS_SPEC_CASE)
    return kNotHandled;
}
}  // namespace bn
