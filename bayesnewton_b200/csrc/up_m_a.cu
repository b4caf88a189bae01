// Fused posterior update instantiations, stationary-kernel group A (see common.cuh).
#include "up_impl.cuh"
namespace bn {
int up_group_m_a(const UpCall& c) {
    BN_GROUP_M_A(BN_UP_SPEC_CASE)
    return kNotHandled;
}
}  // namespace bn
