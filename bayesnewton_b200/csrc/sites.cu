// C ABI of the site kernels (include/bn_b200.h): bn_site_update, bn_expected_density,
// bn_gaussian_expected_log_lik, bn_ep_pseudo_density.  One thread per time step, coalesced
// loads/stores of the per-step scalars, deterministic two-stage reductions for the sums.
#include <cstdlib>
#include <mutex>
#include "sites_impl.cuh"
#include "tma_stage.cuh"

namespace bn {

constexpr int kSiteThreads = 256;
constexpr int kSiteMaxGrid = 148 * 6;  // persistent grid: 6 CTAs per SM, grid-stride over the time steps
// kernels that gather from the probit table keep all of it (128 KB) in shared memory: one CTA of
// 1024 threads per SM
constexpr int kTabThreads = 1024;
constexpr int kTabGrid = 148;
template <bool TAB> constexpr int kNT = TAB ? kTabThreads : kSiteThreads;

// the probit log-density table in device memory, filled once per device (probit_table.cuh)
__device__ __align__(16) double g_probit_tab[kPtDoubles];
static bool g_probit_ready[64] = {false};
static std::mutex g_probit_mutex;

// One upload per device, guarded by a mutex and SYNCHRONOUS: when this returns the table is in device memory for
// every stream and every host thread (an async copy on the first caller's stream would race with a first use
// on another stream).
static int ensure_probit_table(cudaStream_t) {
    int dev = 0;
    BN_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return -1; }
    std::lock_guard<std::mutex> lock(g_probit_mutex);
    if (!g_probit_ready[dev]) {
        const std::vector<double>& t = probit_table_host();
        BN_CUDA(cudaMemcpyToSymbol(g_probit_tab, t.data(), sizeof(double) * kPtDoubles, 0, cudaMemcpyHostToDevice));
        g_probit_ready[dev] = true;
    }
    return 0;
}

// the table's device address for kernels of other translation units (iter_impl.cuh)
int probit_table_device(cudaStream_t st, const double** tab) {
    if (int rc = ensure_probit_table(st)) return rc;
    void* p = nullptr;
    BN_CUDA(cudaGetSymbolAddress(&p, g_probit_tab));
    *tab = (const double*)p;
    return 0;
}

// the cubature-driven probit schemes read the table from shared memory
template <int LIK, int METHOD>
constexpr bool kUsesTable = (LIK == BN_LIK_BERNOULLI_PROBIT) && (METHOD == BN_METHOD_VI || METHOD == BN_METHOD_EP || METHOD == BN_METHOD_PL);

// BN_B200_PROBIT_TABLE=0 in the environment routes the probit schemes through erf()/log() instead
// (validation aid: A/B of the tabulated log-density against libm on the device)
bool probit_table_enabled() {
    static const bool on = [] {
        const char* e = getenv("BN_B200_PROBIT_TABLE");
        return !(e && e[0] == '0');
    }();
    return on;
}

template <bool TAB>
__device__ __forceinline__ const double* stage_table(double* sm) {
    if constexpr (TAB) {
        tma_stage_to_smem(sm, g_probit_tab, (uint32_t)(sizeof(double) * kPtDoubles));
        return sm;
    } else {
        return nullptr;
    }
}

template <int LIK, int METHOD, bool TAB>
__global__ void __launch_bounds__(kNT<TAB>)
site_update_kernel(const __grid_constant__ bn_site_args a, const __grid_constant__ Cub1 cub, const double* cx2,
                   const double* cw2, double* part1, double* part2) {
    extern __shared__ __align__(16) double site_smem[];
    const SiteCtx sc{&cub, stage_table<TAB>(site_smem), cx2, cw2};
    double d1 = 0.0, d2 = 0.0;
    for (long long n = (long long)blockIdx.x * kNT<TAB> + threadIdx.x; n < a.N; n += (long long)gridDim.x * kNT<TAB>) {
        double e1, e2;
        site_update_step<LIK, METHOD, TAB>(a, sc, n, e1, e2);
        d1 += e1;
        d2 += e2;
    }
    if (part1) {
        block_sum_store<kNT<TAB>>(d1, part1);
        block_sum_store<kNT<TAB>>(d2, part2);
    }
}

template <int LIK, int METHOD, bool TAB>
__global__ void __launch_bounds__(kNT<TAB>)
expected_density_kernel(const __grid_constant__ bn_site_args a, const __grid_constant__ Cub1 cub, const double* cx2,
                        const double* cw2, double* values, double* part) {
    extern __shared__ __align__(16) double site_smem[];
    const SiteCtx sc{&cub, stage_table<TAB>(site_smem), cx2, cw2};
    double acc = 0.0;
    for (long long n = (long long)blockIdx.x * kNT<TAB> + threadIdx.x; n < a.N; n += (long long)gridDim.x * kNT<TAB>) {
        double v = expected_density_step<LIK, METHOD, TAB>(a, sc, n);
        if (values) values[n] = v;
        if (!isnan(v)) acc += v;  // nansum
    }
    block_sum_store<kNT<TAB>>(acc, part);
}

// the two per-step sums of a VI / Newton energy in one pass over the posterior marginals: the scheme's likelihood term
// and E_q[log N(pseudo_y | f, pseudo_var)] (utils.py:510-531); the posterior is read once
template <int LIK, int METHOD, bool TAB>
__global__ void __launch_bounds__(kNT<TAB>)
energy_terms_kernel(const __grid_constant__ bn_site_args a, const __grid_constant__ Cub1 cub, const unsigned char* mask,
                    double* part, double* part2) {
    extern __shared__ __align__(16) double site_smem[];
    const SiteCtx sc{&cub, stage_table<TAB>(site_smem), nullptr, nullptr};
    double acc = 0.0, acc2 = 0.0;
    for (long long n = (long long)blockIdx.x * kNT<TAB> + threadIdx.x; n < a.N; n += (long long)gridDim.x * kNT<TAB>) {
        const double v = expected_density_step<LIK, METHOD, TAB>(a, sc, n);
        if (!isnan(v)) acc += v;  // nansum
        acc2 += gaussian_ell_step<1>(a.site_mean, a.post_mean, a.post_cov, a.site_cov, mask, n);
    }
    block_sum_store<kNT<TAB>>(acc, part);
    block_sum_store<kNT<TAB>>(acc2, part2);
}

template <int LIK, int METHOD, bool TAB>
__global__ void __launch_bounds__(kNT<TAB>)
likelihood_stats_kernel(const __grid_constant__ bn_site_args a, const __grid_constant__ Cub1 cub, const double* cx2,
                        const double* cw2, double* val, double* d1, double* d2) {
    extern __shared__ __align__(16) double site_smem[];
    const SiteCtx sc{&cub, stage_table<TAB>(site_smem), cx2, cw2};
    for (long long n = (long long)blockIdx.x * kNT<TAB> + threadIdx.x; n < a.N; n += (long long)gridDim.x * kNT<TAB>)
        likelihood_stats_step<LIK, METHOD, TAB>(a, sc, n, val, d1, d2);
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

// d (likelihood term of energy()) / d (Gaussian variance), summed over the steps: the route from the likelihood's
// hyper-parameter to the energy (the posterior and the sites are StateVars; objax.GradValues(model.energy, model.vars()),
// README.md:56-70).  VI: E_q[log N(y | f, s2)] (likelihoods.py:727-753);  Newton: log N(y | m, s2) (:336-355);
// EP: log N(y | m_c, s2 / a + v_c) + pep_constant(s2, a) at the cavity (:755-782, utils.py:431-445, 534-541).
template <int METHOD>
__global__ void __launch_bounds__(kSiteThreads)
gaussian_param_grad_kernel(const __grid_constant__ bn_site_args a, double* part) {
    const long long n = (long long)blockIdx.x * kSiteThreads + threadIdx.x;
    double v = 0.0;
    if (n < a.N) {
        const double y = a.y[n], s2 = a.lik_param;
        if (!isnan(y)) {
            double m = a.post_mean[n], c = a.post_cov[n];
            if constexpr (METHOD == BN_METHOD_EP) {  // compute_cavity, utils.py:534-541
                const double pn2 = inv1(c + 1e-8);
                c = inv1(pn2 - a.power * a.nat2[n]);
                m = c * (pn2 * m - a.power * a.nat1[n]);
                const double var = s2 / a.power + c, r = y - m;
                v = (-0.5 / var + 0.5 * r * r / (var * var)) / a.power + 0.5 * (1.0 - a.power) / s2;
            } else if constexpr (METHOD == BN_METHOD_VI) {
                const double r = y - m;
                v = -0.5 / s2 + 0.5 * (r * r + c) / (s2 * s2);
            } else {
                const double r = y - m;
                v = -0.5 / s2 + 0.5 * r * r / (s2 * s2);
            }
        }
    }
    block_sum_store<kSiteThreads>(v, part);
}

static int check_site_args(const bn_site_args* a, bool need_y = true) {
    BN_REQUIRE(a != nullptr, "site args are null");
    BN_REQUIRE(a->N >= 0, "N must be non-negative");
    bool het = a->likelihood == BN_LIK_HETEROSCEDASTIC_SOFTPLUS || a->likelihood == BN_LIK_HETEROSCEDASTIC_EXP;
    BN_REQUIRE(a->D == (het ? 2 : 1), "likelihood %d needs D = %d latents, got %d", a->likelihood, het ? 2 : 1, a->D);
    BN_REQUIRE(a->N == 0 || ((a->y || !need_y) && a->post_mean && a->post_cov), "null input array");
    bool closed = (a->likelihood == BN_LIK_GAUSSIAN && (a->method == BN_METHOD_VI || a->method == BN_METHOD_EP)) ||
                  (a->likelihood == BN_LIK_POISSON_EXP && a->method == BN_METHOD_VI);
    if (a->method != BN_METHOD_NEWTON && !closed) {
        BN_REQUIRE(a->Q > 0 && a->cub_x && a->cub_w, "cubature table missing");
        if (!het) BN_REQUIRE(a->Q <= kMaxQ1, "at most %d cubature points are supported for a single latent, got %d", kMaxQ1, a->Q);
    }
    if (a->likelihood == BN_LIK_GAUSSIAN) BN_REQUIRE(a->lik_param > 0.0, "Gaussian variance must be positive");
    if (a->likelihood == BN_LIK_POISSON_EXP) BN_REQUIRE(a->lik_param > 0.0, "Poisson bin size must be positive");
    if (a->likelihood == BN_LIK_STUDENTS_T)
        BN_REQUIRE(a->lik_param > 0.0 && a->lik_param2 > 0.0, "Student-t scale and degrees of freedom must be positive");
    if (a->likelihood == BN_LIK_GAMMA_EXP) BN_REQUIRE(a->lik_param > 0.0, "Gamma shape must be positive");
    if (a->likelihood == BN_LIK_NEGBIN_EXP)
        BN_REQUIRE(a->lik_param > 0.0 && a->lik_param2 > 0.0, "negative-binomial alpha and scale must be positive");
    if (a->likelihood == BN_LIK_BETA_PROBIT) BN_REQUIRE(a->lik_param > 0.0, "Beta scale must be positive");
    if (a->method == BN_METHOD_EP) BN_REQUIRE(a->power > 0.0, "EP power must be positive");
    return 0;
}

// Everything a site launch needs: grid, the 1-D rule by value, the multi-latent rule staged into the
// caller's workspace (after `partials` doubles), dynamic shared memory for the probit table.
struct SitePlan {
    unsigned grid;
    Cub1 cub;
    const double *cx2, *cw2;
    double* partials;
};

static int plan_sites(const bn_site_args* a, size_t partial_doubles, void* workspace, size_t workspace_bytes,
                      cudaStream_t st, SitePlan& p) {
    long long g = (a->N + kSiteThreads - 1) / kSiteThreads;
    p.grid = (unsigned)(g < kSiteMaxGrid ? g : kSiteMaxGrid);
    const bool tab = probit_table_enabled() && a->likelihood == BN_LIK_BERNOULLI_PROBIT &&
                     (a->method == BN_METHOD_VI || a->method == BN_METHOD_EP || a->method == BN_METHOD_PL);
    if (tab) {
        g = (a->N + kTabThreads - 1) / kTabThreads;
        p.grid = (unsigned)(g < kTabGrid ? g : kTabGrid);
    }
    p.cx2 = p.cw2 = nullptr;
    p.partials = (double*)workspace;
    const bool het = a->D == 2;
    const bool has_cub = a->Q > 0 && a->cub_x && a->cub_w;
    make_cub1(het ? 0 : a->Q, has_cub ? a->cub_x : nullptr, has_cub ? a->cub_w : nullptr, p.cub);
    size_t need = partial_doubles * p.grid * sizeof(double);
    if (het && has_cub) need += (size_t)3 * a->Q * sizeof(double);
    BN_REQUIRE(need == 0 || (workspace && workspace_bytes >= need), "workspace too small: need %zu bytes, got %zu", need,
               workspace_bytes);
    if (het && has_cub) {  // the 2-D rule (400 points by default) lives in device memory
        double* t = (double*)workspace + partial_doubles * p.grid;
        BN_CUDA(cudaMemcpyAsync(t, a->cub_x, (size_t)2 * a->Q * sizeof(double), cudaMemcpyHostToDevice, st));
        BN_CUDA(cudaMemcpyAsync(t + 2 * a->Q, a->cub_w, (size_t)a->Q * sizeof(double), cudaMemcpyHostToDevice, st));
        p.cx2 = t;
        p.cw2 = t + 2 * a->Q;
    }
    return 0;
}

template <int LIK, int METHOD>
static size_t table_smem(cudaStream_t st, int& rc) {
    rc = 0;
    if constexpr (!kUsesTable<LIK, METHOD>) return 0;
    if (!probit_table_enabled()) return 0;
    rc = ensure_probit_table(st);
    return sizeof(double) * kPtDoubles;
}

// launch K<L, M, TAB> with TAB decided at run time (only the table-capable pairs instantiate both)
#define BN_SITE_LAUNCH(K, L, M, smem, ...)                                                            \
    do {                                                                                              \
        if (smem) { /* smem != 0 implies kUsesTable<L, M> */                                          \
            BN_CUDA(cudaFuncSetAttribute(K<L, M, kUsesTable<L, M>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            K<L, M, kUsesTable<L, M>><<<p.grid, kNT<kUsesTable<L, M>>, smem, st>>>(__VA_ARGS__);      \
        } else {                                                                                      \
            K<L, M, false><<<p.grid, kSiteThreads, 0, st>>>(__VA_ARGS__);                             \
        }                                                                                             \
    } while (0)

}  // namespace bn

using namespace bn;

extern "C" int bn_site_update(const bn_site_args* a, void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = check_site_args(a)) return rc;
    BN_REQUIRE(a->nat1 && a->nat2, "nat1/nat2 must be given (they are updated in place)");
    if (a->N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    SitePlan p;
    if (int rc = plan_sites(a, a->diffs ? 2 : 0, workspace, workspace_bytes, st, p)) return rc;
    double *p1 = a->diffs ? p.partials : nullptr, *p2 = a->diffs ? p.partials + p.grid : nullptr;
#define X(L, M)                                                                \
    if (a->likelihood == L && a->method == M) {                                \
        int rc;                                                                \
        size_t smem = table_smem<L, M>(st, rc);                                \
        if (rc) return rc;                                                     \
        BN_LAUNCH("site_update", st, BN_SITE_LAUNCH(site_update_kernel, L, M, smem, *a, p.cub, p.cx2, p.cw2, p1, p2)); \
        BN_CUDA(cudaGetLastError());                                           \
        if (a->diffs) {                                                        \
            sum_kernel<false><<<1, 1024, 0, st>>>(p1, p.grid, a->diffs, 1.0 / ((double)a->N * a->D));            \
            sum_kernel<false><<<1, 1024, 0, st>>>(p2, p.grid, a->diffs + 1, 1.0 / ((double)a->N * a->D * a->D)); \
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
    SitePlan p;
    if (int rc = plan_sites(a, 1, workspace, workspace_bytes, st, p)) return rc;
#define X(L, M)                                                                        \
    if (a->likelihood == L && a->method == M) {                                        \
        constexpr int ME = (M == BN_METHOD_PL) ? BN_METHOD_EP : M;                     \
        int rc;                                                                        \
        size_t smem = table_smem<L, ME>(st, rc);                                       \
        if (rc) return rc;                                                             \
        BN_LAUNCH("expected_density", st,                                              \
                  BN_SITE_LAUNCH(expected_density_kernel, L, M, smem, *a, p.cub, p.cx2, p.cw2, values, p.partials)); \
        BN_CUDA(cudaGetLastError());                                                   \
        sum_kernel<false><<<1, 1024, 0, st>>>(p.partials, p.grid, sum, 1.0);           \
        BN_CUDA(cudaGetLastError());                                                   \
        return 0;                                                                      \
    }
    BN_FOR_EACH_SITE(X)
#undef X
    set_error("unsupported (likelihood, method) = (%d, %d)", a->likelihood, a->method);
    return -1;
}

extern "C" int bn_energy_terms(const bn_site_args* a, const uint8_t* mask, double* sums, void* workspace,
                               size_t workspace_bytes, void* stream) {
    if (int rc = check_site_args(a)) return rc;
    BN_REQUIRE(sums != nullptr, "sums output is null");
    BN_REQUIRE(a->D == 1 && (a->method == BN_METHOD_VI || a->method == BN_METHOD_NEWTON),
               "the fused energy terms are built for single-latent VI / Newton (use bn_expected_density + "
               "bn_gaussian_expected_log_lik otherwise)");
    cudaStream_t st = (cudaStream_t)stream;
    if (a->N == 0) { BN_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(double), st)); return 0; }
    BN_REQUIRE(a->site_mean && a->site_cov, "site_mean / site_cov (the pseudo observations) are needed");
    SitePlan p;
    if (int rc = plan_sites(a, 2, workspace, workspace_bytes, st, p)) return rc;
#define X(L, M)                                                                        \
    if (a->likelihood == L && a->method == M && (M == BN_METHOD_VI || M == BN_METHOD_NEWTON) && \
        L != BN_LIK_HETEROSCEDASTIC_SOFTPLUS && L != BN_LIK_HETEROSCEDASTIC_EXP) {     \
        int rc;                                                                        \
        size_t smem = table_smem<L, M>(st, rc);                                        \
        if (rc) return rc;                                                             \
        BN_LAUNCH("energy_terms", st,                                                  \
                  BN_SITE_LAUNCH(energy_terms_kernel, L, M, smem, *a, p.cub, mask, p.partials, p.partials + p.grid)); \
        BN_CUDA(cudaGetLastError());                                                   \
        sum_kernel<false><<<1, 1024, 0, st>>>(p.partials, p.grid, sums, 1.0);          \
        sum_kernel<false><<<1, 1024, 0, st>>>(p.partials + p.grid, p.grid, sums + 1, 1.0); \
        BN_CUDA(cudaGetLastError());                                                   \
        return 0;                                                                      \
    }
    BN_FOR_EACH_SITE(X)
#undef X
    set_error("unsupported (likelihood, method) = (%d, %d)", a->likelihood, a->method);
    return -1;
}

extern "C" int bn_likelihood_param_grad(const bn_site_args* a, double* sum, void* workspace, size_t workspace_bytes,
                                        void* stream) {
    if (int rc = check_site_args(a)) return rc;
    BN_REQUIRE(sum != nullptr, "sum output is null");
    BN_REQUIRE(a->likelihood == BN_LIK_GAUSSIAN, "only the Gaussian likelihood carries a trainable parameter (its variance)");
    BN_REQUIRE(a->method == BN_METHOD_VI || a->method == BN_METHOD_NEWTON || a->method == BN_METHOD_EP,
               "the likelihood-parameter gradient is built for VI, Newton and EP");
    if (a->method == BN_METHOD_EP) BN_REQUIRE(a->nat1 && a->nat2, "EP needs the site natural parameters for the cavity");
    cudaStream_t st = (cudaStream_t)stream;
    if (a->N == 0) { BN_CUDA(cudaMemsetAsync(sum, 0, sizeof(double), st)); return 0; }
    const unsigned grid = (unsigned)((a->N + kSiteThreads - 1) / kSiteThreads);
    BN_REQUIRE(workspace && workspace_bytes >= (size_t)grid * sizeof(double), "workspace too small for %u partials", grid);
    double* part = (double*)workspace;
    if (a->method == BN_METHOD_VI) gaussian_param_grad_kernel<BN_METHOD_VI><<<grid, kSiteThreads, 0, st>>>(*a, part);
    else if (a->method == BN_METHOD_NEWTON) gaussian_param_grad_kernel<BN_METHOD_NEWTON><<<grid, kSiteThreads, 0, st>>>(*a, part);
    else gaussian_param_grad_kernel<BN_METHOD_EP><<<grid, kSiteThreads, 0, st>>>(*a, part);
    BN_CUDA(cudaGetLastError());
    sum_kernel<false><<<1, 1024, 0, st>>>(part, grid, sum, 1.0);
    BN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int bn_likelihood_stats(const bn_site_args* a, double* val, double* d1, double* d2, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    if (int rc = check_site_args(a, a && a->method != BN_METHOD_PL)) return rc;
    if (a->N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    SitePlan p;
    if (int rc = plan_sites(a, 0, workspace, workspace_bytes, st, p)) return rc;
#define X(L, M)                                                                         \
    if (a->likelihood == L && a->method == M) {                                         \
        int rc;                                                                         \
        size_t smem = table_smem<L, M>(st, rc);                                         \
        if (rc) return rc;                                                              \
        BN_SITE_LAUNCH(likelihood_stats_kernel, L, M, smem, *a, p.cub, p.cx2, p.cw2, val, d1, d2); \
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
    // pair sites of the sparse Markov model (2n x 2n blocks, basemodels.py:215-223 through :1021-1031)
    else if (D == 4) BN_LAUNCH("gaussian_ell", st, gaussian_ell_kernel<4><<<grid, kSiteThreads, 0, st>>>(N, pseudo_y, post_mean, post_cov, pseudo_var, mask, values, part));
    else if (D == 6) BN_LAUNCH("gaussian_ell", st, gaussian_ell_kernel<6><<<grid, kSiteThreads, 0, st>>>(N, pseudo_y, post_mean, post_cov, pseudo_var, mask, values, part));
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
