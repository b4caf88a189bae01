// Kalman filter instantiations, array-entry group A (see common.cuh).
#include "filter_impl.cuh"
namespace bn {
int kf_group_a_a(const KfCall& c) {
    BN_GROUP_A_A(BN_KF_ARR_CASE)
    return kNotHandled;
}
}  // namespace bn
