// Kalman filter instantiations, stationary-kernel group A (see common.cuh).
#include "filter_impl.cuh"
namespace bn {
int kf_group_m_a(const KfCall& c) {
    BN_GROUP_M_A(BN_KF_SPEC_CASE)
    return kNotHandled;
}
}  // namespace bn
