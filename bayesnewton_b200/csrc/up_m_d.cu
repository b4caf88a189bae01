// Fused posterior update instantiations, stationary-kernel group D (see common.cuh).
#include "up_impl.cuh"
namespace bn {
int up_group_m_d(const UpCall& c) {
    BN_GROUP_M_D(BN_UP_SPEC_CASE)
    return kNotHandled;
}
}  // namespace bn
