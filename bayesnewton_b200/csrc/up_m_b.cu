// Fused posterior update instantiations, stationary-kernel group B (see common.cuh).
#include "up_impl.cuh"
namespace bn {
int up_group_m_b(const UpCall& c) {
    BN_GROUP_M_B(BN_UP_SPEC_CASE)
    return kNotHandled;
}
}  // namespace bn
