// C ABI of the fused posterior update (include/bn_b200.h): bn_update_posterior and the three
// phases of its time-sharded form.  Kernels: up_impl.cuh, instantiated per group in up_m_*.cu.
#include "up_impl.cuh"

namespace bn {
int up_group_m_a(const UpCall&);
int up_group_m_b(const UpCall&);
int up_group_m_c(const UpCall&);
int up_group_m_d(const UpCall&);

static int up_dispatch(const UpCall& c) {
    int r;
    if ((r = up_group_m_a(c)) != kNotHandled) return r;
    if ((r = up_group_m_b(c)) != kNotHandled) return r;
    if ((r = up_group_m_c(c)) != kNotHandled) return r;
    if ((r = up_group_m_d(c)) != kNotHandled) return r;
    set_error("unsupported kernel spec: family %d with %d components", c.spec->family, c.spec->n_components);
    return -1;
}

static int up_check(const bn_kernel_spec* k, int64_t N, const double* dt, const double* y, const double* R) {
    BN_REQUIRE(k != nullptr, "kernel spec is null");
    BN_REQUIRE(N >= 0, "N must be non-negative");
    BN_REQUIRE(N == 0 || (dt && y && R), "null input array");
    return 0;
}
}  // namespace bn

using namespace bn;

extern "C" size_t bn_update_posterior_workspace_bytes(const bn_kernel_spec* k, int64_t N) {
    int d = bn_state_dim(k);
    if (d < 1 || N < 0) return 0;
    size_t doubles = 0;
    switch (d) {
        case 1: doubles = up_ws_doubles<1>(N); break;
        case 2: doubles = up_ws_doubles<2>(N); break;
        case 3: doubles = up_ws_doubles<3>(N); break;
        case 4: doubles = up_ws_doubles<4>(N); break;
        case 6: doubles = up_ws_doubles<6>(N); break;
        default: return 0;
    }
    return (doubles + 64) * sizeof(double);
}

extern "C" int bn_update_posterior(const bn_kernel_spec* k, int64_t N, const double* dt, const double* pseudo_y,
                                   const double* pseudo_var, const uint8_t* mask, double* ell, double* post_mean,
                                   double* post_cov, void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = up_check(k, N, dt, pseudo_y, pseudo_var)) return rc;
    BN_REQUIRE(N == 0 || (post_mean && post_cov), "null output array");
    UpCall c{k, UpIO{N, dt, pseudo_y, pseudo_var, mask, post_mean, post_cov}, ell, workspace, workspace_bytes,
             (cudaStream_t)stream, UP_ALL, 0, 1, nullptr, nullptr};
    return up_dispatch(c);
}

extern "C" int bn_update_posterior_grad(const bn_kernel_spec* k, int64_t N, const double* dt, const double* pseudo_y,
                                        const double* pseudo_var, double* ell, double* post_mean, double* post_cov,
                                        double* dell_dvariance, double* dell_dlengthscale, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    if (int rc = up_check(k, N, dt, pseudo_y, pseudo_var)) return rc;
    BN_REQUIRE(N > 0, "the hyper-gradient needs at least one step");
    BN_REQUIRE(post_mean && post_cov && dell_dvariance && dell_dlengthscale, "null output array");
    UpCall c{k, UpIO{N, dt, pseudo_y, pseudo_var, nullptr, post_mean, post_cov}, ell, workspace, workspace_bytes,
             (cudaStream_t)stream, UP_ALL, 0, 1, nullptr, nullptr, 1, dell_dvariance, dell_dlengthscale};
    return up_dispatch(c);
}

extern "C" int bn_up_shard_reduce(const bn_kernel_spec* k, int64_t N, int rank, int world, const double* dt,
                                  const double* pseudo_y, const double* pseudo_var, double* carry, int want_grad,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = up_check(k, N, dt, pseudo_y, pseudo_var)) return rc;
    BN_REQUIRE(N > 0, "a time shard must hold at least one step");
    BN_REQUIRE(rank >= 0 && rank < world, "rank %d outside world %d", rank, world);
    BN_REQUIRE(carry != nullptr, "carry output is null");
    UpCall c{k, UpIO{N, dt, pseudo_y, pseudo_var, nullptr, nullptr, nullptr}, nullptr, workspace, workspace_bytes,
             (cudaStream_t)stream, UP_REDUCE, rank, world, carry, nullptr, want_grad != 0, nullptr, nullptr};
    return up_dispatch(c);
}

extern "C" int bn_up_shard_filter(const bn_kernel_spec* k, int64_t N, int rank, int world, const double* kf_carries,
                                  const double* dt, const double* pseudo_y, const double* pseudo_var,
                                  const uint8_t* mask, double* ell, double* rts_carry, int want_grad,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = up_check(k, N, dt, pseudo_y, pseudo_var)) return rc;
    BN_REQUIRE(N > 0, "a time shard must hold at least one step");
    BN_REQUIRE(rank >= 0 && rank < world, "rank %d outside world %d", rank, world);
    BN_REQUIRE(kf_carries && rts_carry, "null carry array");
    BN_REQUIRE(!(want_grad && mask), "the hyper-gradient is not available with a mask (see bn_update_posterior_grad)");
    UpCall c{k, UpIO{N, dt, pseudo_y, pseudo_var, mask, nullptr, nullptr}, ell, workspace, workspace_bytes,
             (cudaStream_t)stream, UP_FILTER, rank, world, rts_carry, kf_carries, want_grad != 0, nullptr, nullptr};
    return up_dispatch(c);
}

extern "C" int bn_up_shard_smooth(const bn_kernel_spec* k, int64_t N, int rank, int world, const double* rts_carries,
                                  const double* dt, double* post_mean, double* post_cov, double* dell_dvariance,
                                  double* dell_dlengthscale, void* workspace, size_t workspace_bytes, void* stream) {
    BN_REQUIRE(k != nullptr, "kernel spec is null");
    BN_REQUIRE(N > 0, "a time shard must hold at least one step");
    BN_REQUIRE(rank >= 0 && rank < world, "rank %d outside world %d", rank, world);
    BN_REQUIRE(rts_carries && dt && post_mean && post_cov, "null array");
    BN_REQUIRE((dell_dvariance == nullptr) == (dell_dlengthscale == nullptr), "give both gradient outputs or neither");
    UpCall c{k, UpIO{N, dt, nullptr, nullptr, nullptr, post_mean, post_cov}, nullptr, workspace, workspace_bytes,
             (cudaStream_t)stream, UP_SMOOTH, rank, world, nullptr, rts_carries, dell_dvariance != nullptr,
             dell_dvariance, dell_dlengthscale};
    return up_dispatch(c);
}
