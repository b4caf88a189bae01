// RTS smoother instantiations, stationary-kernel group B (see common.cuh).
#include "smoother_impl.cuh"
namespace bn {
int rts_group_m_b(const RtsCall& c) {
    BN_GROUP_M_B(BN_RTS_SPEC_CASE)
    return kNotHandled;
}
}  // namespace bn
