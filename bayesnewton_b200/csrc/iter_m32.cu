// Fused-iteration instantiations for BN_MATERN32 (one component); see iter_impl.cuh.
#include "iter_impl.cuh"
namespace BN_NS {
int it_group_m32(const ItCall& c) {
    if (c.spec->family == BN_MATERN32 && c.spec->n_components == 1) return it_run<FastGen<BN_MATERN32, 1>>(c);
    return kNotHandled;
}
}  // namespace BN_NS
