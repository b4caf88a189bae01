// Kalman filter instantiations, stationary-kernel group D (see common.cuh).
#include "filter_impl.cuh"
namespace bn {
int kf_group_m_d(const KfCall& c) {
    BN_GROUP_M_D(BN_KF_SPEC_CASE)
    return kNotHandled;
}
}  // namespace bn
