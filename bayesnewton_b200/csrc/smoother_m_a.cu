// RTS smoother instantiations, stationary-kernel group A (see common.cuh).
#include "smoother_impl.cuh"
namespace bn {
int rts_group_m_a(const RtsCall& c) {
    BN_GROUP_M_A(BN_RTS_SPEC_CASE)
    return kNotHandled;
}
}  // namespace bn
