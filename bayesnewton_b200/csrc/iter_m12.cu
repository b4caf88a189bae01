// Fused-iteration instantiations for BN_MATERN12 (one component); see iter_impl.cuh.
#include "iter_impl.cuh"
namespace BN_NS {
int it_group_m12(const ItCall& c) {
    if (c.spec->family == BN_MATERN12 && c.spec->n_components == 1) return it_run<FastGen<BN_MATERN12, 1>>(c);
    return kNotHandled;
}
}  // namespace BN_NS
