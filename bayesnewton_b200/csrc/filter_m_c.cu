// Kalman filter instantiations, stationary-kernel group C (see common.cuh).
#include "filter_impl.cuh"
namespace bn {
int kf_group_m_c(const KfCall& c) {
    BN_GROUP_M_C(BN_KF_SPEC_CASE)
    return kNotHandled;
}
}  // namespace bn
