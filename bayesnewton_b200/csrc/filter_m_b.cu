// Kalman filter instantiations, stationary-kernel group B (see common.cuh).
#include "filter_impl.cuh"
namespace bn {
int kf_group_m_b(const KfCall& c) {
    BN_GROUP_M_B(BN_KF_SPEC_CASE)
    return kNotHandled;
}
}  // namespace bn
