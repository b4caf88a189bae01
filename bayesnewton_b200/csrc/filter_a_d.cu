// Kalman filter instantiations, array-entry group D: the pairs filter of the sparse Markov model (see common.cuh).
#include "filter_impl.cuh"
namespace bn {
int kf_group_a_d(const KfCall& c) {
    BN_GROUP_A_D(BN_KF_ARR_CASE)
    return kNotHandled;
}
}  // namespace bn
