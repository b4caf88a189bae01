// C ABI of the site kernels (include/bn_b200.h): bn_site_update, bn_expected_density,
// bn_gaussian_expected_log_lik, bn_ep_pseudo_density.  One thread per time step, coalesced
// loads/stores of the per-step scalars, deterministic two-stage reductions for the sums.
#include "sites_impl.cuh"

namespace bn {

constexpr int kSiteThreads = 256;

template <int LIK, int METHOD>
__global__ void __launch_bounds__(kSiteThreads) site_update_kernel(bn_site_args a, double* part1, double* part2) {
    const long long n = (long long)blockIdx.x * kSiteThreads + threadIdx.x;
    double d1 = 0.0, d2 = 0.0;
    if (n < a.N) site_update_step<LIK, METHOD>(a, n, d1, d2);
    if (part1) {
        block_sum_store<kSiteThreads>(d1, part1);
        block_sum_store<kSiteThreads>(d2, part2);
    }
}

template <int LIK, int METHOD>
__global__ void __launch_bounds__(kSiteThreads) expected_density_kernel(bn_site_args a, double* values, double* part) {
    const long long n = (long long)blockIdx.x * kSiteThreads + threadIdx.x;
    double v = 0.0;
    if (n < a.N) {
        v = expected_density_step<LIK, METHOD>(a, n);
        if (values) values[n] = v;
        if (isnan(v)) v = 0.0;  // nansum
    }
    block_sum_store<kSiteThreads>(v, part);
}

template <int LIK, int METHOD>
__global__ void __launch_bounds__(kSiteThreads)
likelihood_stats_kernel(bn_site_args a, double* val, double* d1, double* d2) {
    const long long n = (long long)blockIdx.x * kSiteThreads + threadIdx.x;
    if (n < a.N) likelihood_stats_step<LIK, METHOD>(a, n, val, d1, d2);
}

template <int D>
__global__ void __launch_bounds__(kSiteThreads)
gaussian_ell_kernel(long long N, const double* py, const double* pm, const double* pV, const double* pR,
                    const unsigned char* mask, double* values, double* part) {
    const long long n = (long long)blockIdx.x * kSiteThreads + threadIdx.x;
    double v = 0.0;
    if (n < N) {
        v = gaussian_ell_step<D>(py, pm, pV, pR, mask, n);
        if (values) values[n] = v;
    }
    block_sum_store<kSiteThreads>(v, part);
}

template <int D>
__global__ void __launch_bounds__(kSiteThreads)
ep_pseudo_kernel(long long N, double power, int with_const, const double* py, const double* pR, const double* pm,
                 const double* pV, const double* n1, const double* n2, const unsigned char* mask, double* part) {
    const long long n = (long long)blockIdx.x * kSiteThreads + threadIdx.x;
    double v = 0.0;
    if (n < N) {
        v = ep_pseudo_step<D>(power, with_const, py, pR, pm, pV, n1, n2, mask, n);
        if (isnan(v)) v = 0.0;  // nansum (inference.py:321)
    }
    block_sum_store<kSiteThreads>(v, part);
}

static int check_site_args(const bn_site_args* a, bool need_y = true) {
    BN_REQUIRE(a != nullptr, "site args are null");
    BN_REQUIRE(a->N >= 0, "N must be non-negative");
    bool het = a->likelihood == BN_LIK_HETEROSCEDASTIC_SOFTPLUS || a->likelihood == BN_LIK_HETEROSCEDASTIC_EXP;
    BN_REQUIRE(a->D == (het ? 2 : 1), "likelihood %d needs D = %d latents, got %d", a->likelihood, het ? 2 : 1, a->D);
    BN_REQUIRE(a->N == 0 || ((a->y || !need_y) && a->post_mean && a->post_cov), "null input array");
    bool closed = a->likelihood == BN_LIK_GAUSSIAN && (a->method == BN_METHOD_VI || a->method == BN_METHOD_EP);
    if (a->method != BN_METHOD_NEWTON && !closed)
        BN_REQUIRE(a->Q > 0 && a->cub_x && a->cub_w, "cubature table missing");
    if (a->likelihood == BN_LIK_GAUSSIAN) BN_REQUIRE(a->lik_param > 0.0, "Gaussian variance must be positive");
    if (a->method == BN_METHOD_EP) BN_REQUIRE(a->power > 0.0, "EP power must be positive");
    return 0;
}

}  // namespace bn

using namespace bn;

extern "C" int bn_site_update(const bn_site_args* a, void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = check_site_args(a)) return rc;
    BN_REQUIRE(a->nat1 && a->nat2, "nat1/nat2 must be given (they are updated in place)");
    if (a->N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = (unsigned)((a->N + kSiteThreads - 1) / kSiteThreads);
    double *p1 = nullptr, *p2 = nullptr;
    if (a->diffs) {
        BN_REQUIRE(workspace && workspace_bytes >= 2ull * grid * sizeof(double), "workspace too small for %u partials",
                   grid);
        p1 = (double*)workspace;
        p2 = p1 + grid;
    }
#define X(L, M)                                                                \
    if (a->likelihood == L && a->method == M) {                                \
        BN_LAUNCH("site_update", st, (site_update_kernel<L, M><<<grid, kSiteThreads, 0, st>>>(*a, p1, p2))); \
        BN_CUDA(cudaGetLastError());                                           \
        if (a->diffs) {                                                        \
            sum_kernel<false><<<1, 1024, 0, st>>>(p1, grid, a->diffs, 1.0 / ((double)a->N * a->D));            \
            sum_kernel<false><<<1, 1024, 0, st>>>(p2, grid, a->diffs + 1, 1.0 / ((double)a->N * a->D * a->D)); \
            BN_CUDA(cudaGetLastError());                                       \
        }                                                                      \
        return 0;                                                              \
    }
    BN_FOR_EACH_SITE(X)
#undef X
    set_error("unsupported (likelihood, method) = (%d, %d)", a->likelihood, a->method);
    return -1;
}

extern "C" int bn_expected_density(const bn_site_args* a, double* values, double* sum, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    if (int rc = check_site_args(a)) return rc;
    BN_REQUIRE(sum != nullptr, "sum output is null");
    if (a->method == BN_METHOD_EP || a->method == BN_METHOD_PL)
        BN_REQUIRE(a->nat1 && a->nat2, "EP/PL energies need the site natural parameters for the cavity");
    cudaStream_t st = (cudaStream_t)stream;
    if (a->N == 0) { BN_CUDA(cudaMemsetAsync(sum, 0, sizeof(double), st)); return 0; }
    unsigned grid = (unsigned)((a->N + kSiteThreads - 1) / kSiteThreads);
    BN_REQUIRE(workspace && workspace_bytes >= (size_t)grid * sizeof(double), "workspace too small for %u partials", grid);
    double* part = (double*)workspace;
#define X(L, M)                                                                        \
    if (a->likelihood == L && a->method == M) {                                        \
        BN_LAUNCH("expected_density", st,                                              \
                  (expected_density_kernel<L, M><<<grid, kSiteThreads, 0, st>>>(*a, values, part))); \
        BN_CUDA(cudaGetLastError());                                                   \
        sum_kernel<false><<<1, 1024, 0, st>>>(part, grid, sum, 1.0);                   \
        BN_CUDA(cudaGetLastError());                                                   \
        return 0;                                                                      \
    }
    BN_FOR_EACH_SITE(X)
#undef X
    set_error("unsupported (likelihood, method) = (%d, %d)", a->likelihood, a->method);
    return -1;
}

extern "C" int bn_likelihood_stats(const bn_site_args* a, double* val, double* d1, double* d2, void* stream) {
    if (int rc = check_site_args(a, a && a->method != BN_METHOD_PL)) return rc;
    if (a->N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = (unsigned)((a->N + kSiteThreads - 1) / kSiteThreads);
#define X(L, M)                                                                         \
    if (a->likelihood == L && a->method == M) {                                         \
        likelihood_stats_kernel<L, M><<<grid, kSiteThreads, 0, st>>>(*a, val, d1, d2);  \
        BN_CUDA(cudaGetLastError());                                                    \
        return 0;                                                                       \
    }
    BN_FOR_EACH_SITE(X)
#undef X
    set_error("unsupported (likelihood, method) = (%d, %d)", a->likelihood, a->method);
    return -1;
}

extern "C" int bn_gaussian_expected_log_lik(int64_t N, int D, const double* pseudo_y, const double* post_mean,
                                            const double* post_cov, const double* pseudo_var, const uint8_t* mask,
                                            double* values, double* sum, void* workspace, size_t workspace_bytes,
                                            void* stream) {
    BN_REQUIRE(N >= 0, "N must be non-negative");
    BN_REQUIRE(sum != nullptr, "sum output is null");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) { BN_CUDA(cudaMemsetAsync(sum, 0, sizeof(double), st)); return 0; }
    BN_REQUIRE(pseudo_y && post_mean && post_cov && pseudo_var, "null input array");
    unsigned grid = (unsigned)((N + kSiteThreads - 1) / kSiteThreads);
    BN_REQUIRE(workspace && workspace_bytes >= (size_t)grid * sizeof(double), "workspace too small for %u partials", grid);
    double* part = (double*)workspace;
    if (D == 1) BN_LAUNCH("gaussian_ell", st, gaussian_ell_kernel<1><<<grid, kSiteThreads, 0, st>>>(N, pseudo_y, post_mean, post_cov, pseudo_var, mask, values, part));
    else if (D == 2) gaussian_ell_kernel<2><<<grid, kSiteThreads, 0, st>>>(N, pseudo_y, post_mean, post_cov, pseudo_var, mask, values, part);
    else if (D == 3) gaussian_ell_kernel<3><<<grid, kSiteThreads, 0, st>>>(N, pseudo_y, post_mean, post_cov, pseudo_var, mask, values, part);
    else { set_error("unsupported site dimension %d", D); return -1; }
    BN_CUDA(cudaGetLastError());
    sum_kernel<false><<<1, 1024, 0, st>>>(part, grid, sum, 1.0);
    BN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int bn_ep_pseudo_density(int64_t N, int D, double power, int with_pep_constant, const double* pseudo_y,
                                    const double* pseudo_var, const double* post_mean, const double* post_cov,
                                    const double* nat1, const double* nat2, const uint8_t* mask, double* sum,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    BN_REQUIRE(N >= 0, "N must be non-negative");
    BN_REQUIRE(sum != nullptr, "sum output is null");
    BN_REQUIRE(power > 0.0, "EP power must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) { BN_CUDA(cudaMemsetAsync(sum, 0, sizeof(double), st)); return 0; }
    BN_REQUIRE(pseudo_y && pseudo_var && post_mean && post_cov && nat1 && nat2, "null input array");
    unsigned grid = (unsigned)((N + kSiteThreads - 1) / kSiteThreads);
    BN_REQUIRE(workspace && workspace_bytes >= (size_t)grid * sizeof(double), "workspace too small for %u partials", grid);
    double* part = (double*)workspace;
    if (D == 1) ep_pseudo_kernel<1><<<grid, kSiteThreads, 0, st>>>(N, power, with_pep_constant, pseudo_y, pseudo_var, post_mean, post_cov, nat1, nat2, mask, part);
    else if (D == 2) ep_pseudo_kernel<2><<<grid, kSiteThreads, 0, st>>>(N, power, with_pep_constant, pseudo_y, pseudo_var, post_mean, post_cov, nat1, nat2, mask, part);
    else { set_error("unsupported site dimension %d", D); return -1; }
    BN_CUDA(cudaGetLastError());
    sum_kernel<false><<<1, 1024, 0, st>>>(part, grid, sum, 1.0);
    BN_CUDA(cudaGetLastError());
    return 0;
}
