// Kalman filter instantiations, array-entry group B (see common.cuh).
#include "filter_impl.cuh"
namespace bn {
int kf_group_a_b(const KfCall& c) {
    BN_GROUP_A_B(BN_KF_ARR_CASE)
    return kNotHandled;
}
}  // namespace bn
