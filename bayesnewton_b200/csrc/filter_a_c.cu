// Kalman filter instantiations, array-entry group C (see common.cuh).
#include "filter_impl.cuh"
namespace bn {
int kf_group_a_c(const KfCall& c) {
    BN_GROUP_A_C(BN_KF_ARR_CASE)
    return kNotHandled;
}
}  // namespace bn
