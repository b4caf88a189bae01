// Fused posterior update instantiations, stationary-kernel group C (see common.cuh).
#include "up_impl.cuh"
namespace bn {
int up_group_m_c(const UpCall& c) {
    BN_GROUP_M_C(BN_UP_SPEC_CASE)
    return kNotHandled;
}
}  // namespace bn
