// Sparse Markov GP: the pieces around the pairs filter.
//   bn_pairs_discretise       construct_pair of kalman_filter_pairs (bayesnewton/ops.py:411-419): A_pair, Q_pair per
//                             transition; the filter itself is the array-level entry with (d, D) = (2n, 2n), H = I
//   bn_build_joint            vmap(build_joint) over all transitions (utils.py:544-553, basemodels.py:996-1006)
//   bn_sparse_site_update     the data pass of a VI iteration of SparseMarkovGaussianProcess: per data point
//                             compute_conditional_statistics (utils.py:173-215), conditional_posterior_to_data
//                             (basemodels.py:1071-1104), the likelihood's variational expectation, ensure_psd,
//                             conditional_data_to_posterior (:1106-1112), newton_update (inference.py:21-39); per
//                             transition group_natural_params (:1114-1138; utils.py:218-224 is a sequential
//                             scatter-add lax.scan there), the damped update (inference.py:83-86) and reparametrise
//                             (basemodels.py:85-100) -- ONE kernel, one warp per transition, data read once.
//   bn_sparse_expected_density   sum_n E_q[log p(y_n | f_n)] through the same conditional (inference.py:197-222)
// The inputs are sorted, so the data of one transition are a contiguous range [start[m], start[m+1]): the
// scatter-add becomes a segmented sum with a fixed reduction order (run-to-run bit-stable, no atomics).
#include "common.cuh"
#include "gen.cuh"
#include "condstats.cuh"
#include "sites_impl.cuh"

namespace bn {

template <int FAM>
__global__ void pairs_discretise_kernel(MaternGen<FAM, 1> gen, long long Mt, const double* dz, double* Ap, double* Qp) {
    constexpr int n = FamilyDim<FAM>::value, p = 2 * n;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Mt) return;
    double A[n * n], Q[symn(n)];
    gen.step_dt(dz[k], A, Q);
    double* a = Ap + k * p * p;
    double* q = Qp + k * p * p;
#pragma unroll
    for (int i = 0; i < p; ++i)
#pragma unroll
        for (int j = 0; j < p; ++j) {
            double av = 0.0, qv = 0.0;
            if (i < n && j >= n) av = (j - n == i) ? 1.0 : 0.0;            // [[0, I], [0, A]]
            if (i >= n && j >= n) { av = A[(i - n) * n + (j - n)]; qv = Q[sidx(i - n, j - n)]; }
            if (i < n && j < n) qv = (i == j) ? 1e-32 : 0.0;              // jitter block (ops.py:415)
            a[i * p + j] = av;
            q[i * p + j] = qv;
        }
}

template <int FAM>
__global__ void build_joint_kernel(MaternGen<FAM, 1> gen, long long Mt, const double* sm, const double* sP, const double* gain,
                                   double* pm, double* pV) {
    constexpr int n = FamilyDim<FAM>::value, p = 2 * n;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= Mt) return;
    // augmented arrays: index 0 and Mt are the dummy states (minf, Pinf); gains carry a leading zero
    double Pinf[symn(n)];
    gen.pinf(Pinf);
    double ml[n], mr[n], Cl[n * n], Cr[n * n], G[n * n];
#pragma unroll
    for (int i = 0; i < n; ++i) {
        ml[i] = m == 0 ? 0.0 : sm[(m - 1) * n + i];
        mr[i] = m == Mt - 1 ? 0.0 : sm[m * n + i];
#pragma unroll
        for (int j = 0; j < n; ++j) {
            Cl[i * n + j] = m == 0 ? Pinf[sidx(i, j)] : sP[(m - 1) * n * n + i * n + j];
            Cr[i * n + j] = m == Mt - 1 ? Pinf[sidx(i, j)] : sP[m * n * n + i * n + j];
            G[i * n + j] = m == 0 ? 0.0 : gain[(m - 1) * n * n + i * n + j];
        }
    }
    double X[n * n];
    matmul<n, n, n>(G, Cr, X);
#pragma unroll
    for (int i = 0; i < n; ++i) {
        pm[m * p + i] = ml[i];
        pm[m * p + n + i] = mr[i];
#pragma unroll
        for (int j = 0; j < n; ++j) {
            pV[m * p * p + i * p + j] = Cl[i * n + j];
            pV[m * p * p + i * p + n + j] = X[i * n + j];
            pV[m * p * p + (n + i) * p + j] = X[j * n + i];
            pV[m * p * p + (n + i) * p + n + j] = Cr[i * n + j];
        }
    }
}

struct SparseArgs {
    long long N, Mz;          // data points, inducing points (Mt = Mz + 1 transitions)
    const double* x;          // [N] sorted
    const double* y;          // [N]
    const double* z;          // [Mz] sorted
    const long long* start;   // [Mt + 1] first data index of each transition
    const double* pm;         // [Mt, 2n]      joint posterior mean of neighbouring inducing states
    const double* pV;         // [Mt, 2n, 2n]
    double lik_param;
    double lr;
    int ensure_psd;
    double* nat1;             // [Mt, 2n]      in/out
    double* nat2;             // [Mt, 2n, 2n]  in/out
    double* site_mean;        // [Mt, 2n]
    double* site_cov;         // [Mt, 2n, 2n]
    double* partials;         // [3, Mt]  per-transition partial sums: |d nat1|, |d nat2|, expected density
};

// mean_f, var_f of data point n and the projection w = H [P1, W] (1 x 2n)
template <int FAM>
BN_DEV void sparse_point(const MaternGen<FAM, 1>& gen, const SparseArgs& a, long long m, long long i, const double* pmv,
                         const double* pVv, double* w, double& mean_f, double& var_f) {
    constexpr int n = FamilyDim<FAM>::value, p = 2 * n;
    const long long Mt = a.Mz + 1;
    const double xt = a.x[i];
    const double xl = m == 0 ? -1e10 : a.z[m - 1];
    const double xr = m == Mt - 1 ? 1e10 : a.z[m];
    double P1[n * n], W[n * n], Tm[n * n];
    cond_stats(gen, xt - xl, xr - xt, P1, W, Tm);
#pragma unroll
    for (int j = 0; j < n; ++j) { w[j] = P1[j]; w[n + j] = W[j]; }
    double mf = 0.0, vf = Tm[0];
#pragma unroll
    for (int r = 0; r < p; ++r) {
        mf = fma(w[r], pmv[r], mf);
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < p; ++c) t = fma(pVv[r * p + c], w[c], t);
        vf = fma(w[r], t, vf);
    }
    mean_f = mf;
    var_f = vf;
}

template <int FAM, int LIK, bool UPDATE>
__global__ void __launch_bounds__(128) sparse_site_kernel(MaternGen<FAM, 1> gen, SparseArgs a, Cub1 cub) {
    constexpr int n = FamilyDim<FAM>::value, p = 2 * n, ps = symn(p);
    const long long Mt = a.Mz + 1;
    const int lane = threadIdx.x & 31;
    const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= Mt) return;
    Lik1<LIK, false> lik{a.lik_param, nullptr};
    double pmv[p], pVv[p * p];
#pragma unroll
    for (int r = 0; r < p; ++r) pmv[r] = a.pm[m * p + r];
#pragma unroll
    for (int r = 0; r < p * p; ++r) pVv[r] = a.pV[m * p * p + r];
    double s1[p], s2[ps], val = 0.0;
#pragma unroll
    for (int r = 0; r < p; ++r) s1[r] = 0.0;
#pragma unroll
    for (int r = 0; r < ps; ++r) s2[r] = 0.0;
    const long long i0 = a.start[m], i1 = a.start[m + 1];
    for (long long i = i0 + lane; i < i1; i += 32) {
        double w[p], mean_f, var_f;
        sparse_point<FAM>(gen, a, m, i, pmv, pVv, w, mean_f, var_f);
        SiteStats1 st = site_stats_1<LIK, BN_METHOD_VI, true, false>(lik, a.y[i], mean_f, var_f, 0.0, 0.0, 1.0, cub);
        if (!isnan(st.val)) val += st.val;  // nansum (inference.py:218)
        if constexpr (UPDATE) {
            double h = a.ensure_psd ? ensure_psd1(st.hess) : st.hess;
            // conditional_data_to_posterior + newton_update: hessian W^T h W (NaN -> -1e-6 entry-wise),
            // jacobian W^T j (NaN -> hessian @ mean);  nat1 = jac - hess mean,  nat2 = -hess
            const bool hn = isnan(h), jn = isnan(st.jac);
            double Hm[p];  // hess @ mean
            double wm = 0.0, sm_ = 0.0;
#pragma unroll
            for (int r = 0; r < p; ++r) { wm = fma(w[r], pmv[r], wm); sm_ += pmv[r]; }
#pragma unroll
            for (int r = 0; r < p; ++r) Hm[r] = hn ? -1e-6 * sm_ : h * w[r] * wm;
#pragma unroll
            for (int r = 0; r < p; ++r) {
                const double jq = jn ? Hm[r] : w[r] * st.jac;
                s1[r] += jq - Hm[r];
#pragma unroll
                for (int c = 0; c <= r; ++c) s2[sidx(r, c)] += hn ? 1e-6 : -(h * w[r] * w[c]);
            }
        }
    }
    // fixed-order warp reduction
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        val += __shfl_xor_sync(0xffffffffu, val, off);
        if constexpr (UPDATE) {
#pragma unroll
            for (int r = 0; r < p; ++r) s1[r] += __shfl_xor_sync(0xffffffffu, s1[r], off);
#pragma unroll
            for (int r = 0; r < ps; ++r) s2[r] += __shfl_xor_sync(0xffffffffu, s2[r], off);
        }
    }
    if (lane != 0) return;
    a.partials[2 * Mt + m] = val;
    if constexpr (UPDATE) {
        const double cnt = (double)(i1 - i0);
        const double frac = 1.0 - cnt / fmax(cnt, 1.0);  // full batch: counter = num_neighbours (basemodels.py:1132-1135)
        double d1 = 0.0, d2 = 0.0;
        double n1[p], n2[ps];
#pragma unroll
        for (int r = 0; r < p; ++r) {
            const double old = a.nat1[m * p + r];
            const double nw = s1[r] + frac * old;
            d1 += fabs(nw - old);
            n1[r] = (1.0 - a.lr) * old + a.lr * nw;
            a.nat1[m * p + r] = n1[r];
        }
#pragma unroll
        for (int r = 0; r < p; ++r)
#pragma unroll
            for (int c = 0; c < p; ++c) {
                const double old = a.nat2[m * p * p + r * p + c];
                const double nw = s2[sidx(r, c)] + frac * old + (r == c ? 1e-8 : 0.0);
                d2 += fabs(nw - old);
                const double v = (1.0 - a.lr) * old + a.lr * nw;
                if (c <= r) n2[sidx(r, c)] = v;
                a.nat2[m * p * p + r * p + c] = v;
            }
        a.partials[m] = d1;
        a.partials[Mt + m] = d2;
        // reparametrise (basemodels.py:85-90): cov = nat2^-1, mean = cov nat1
        double Ci[ps];
        sym_inverse<p>(n2, Ci);
#pragma unroll
        for (int r = 0; r < p; ++r) {
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < p; ++c) {
                t = fma(Ci[sidx(r, c)], n1[c], t);
                a.site_cov[m * p * p + r * p + c] = Ci[sidx(r, c)];
            }
            a.site_mean[m * p + r] = t;
        }
    }
}

}  // namespace bn

using namespace bn;

#define SPARSE_FAMILIES(X) X(BN_MATERN12) X(BN_MATERN32) X(BN_MATERN52)
#define SPARSE_LIKS(X, F) X(F, BN_LIK_GAUSSIAN) X(F, BN_LIK_BERNOULLI_PROBIT) X(F, BN_LIK_BERNOULLI_LOGIT) X(F, BN_LIK_POISSON_EXP)

static int sparse_check_spec(const bn_kernel_spec* k) {
    BN_REQUIRE(k != nullptr, "kernel spec is null");
    BN_REQUIRE(k->n_components == 1, "the sparse Markov path takes a single-component kernel");
    BN_REQUIRE(k->family == BN_MATERN12 || k->family == BN_MATERN32 || k->family == BN_MATERN52,
               "kernel family %d not available on the sparse Markov path (pairs of Matern-1/2, -3/2, -5/2 states)", k->family);
    return 0;
}

extern "C" int bn_pairs_discretise(const bn_kernel_spec* k, int64_t Mt, const double* dz, double* Apairs, double* Qpairs,
                                   void* stream) {
    if (int rc = sparse_check_spec(k)) return rc;
    BN_REQUIRE(Mt >= 0, "Mt must be non-negative");
    if (Mt == 0) return 0;
    BN_REQUIRE(dz && Apairs && Qpairs, "null array");
    const unsigned grid = (unsigned)((Mt + 127) / 128);
#define X(FAM)                                                                                                    \
    if (k->family == FAM) {                                                                                       \
        MaternGen<FAM, 1> gen; gen.spec = *k; gen.dt = nullptr;                                                   \
        BN_LAUNCH("pairs_discretise", (cudaStream_t)stream,                                                       \
                  pairs_discretise_kernel<FAM><<<grid, 128, 0, (cudaStream_t)stream>>>(gen, Mt, dz, Apairs, Qpairs)); \
    }
    SPARSE_FAMILIES(X)
#undef X
    BN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int bn_build_joint(const bn_kernel_spec* k, int64_t Mt, const double* mean, const double* cov, const double* gain,
                              double* joint_mean, double* joint_cov, void* stream) {
    if (int rc = sparse_check_spec(k)) return rc;
    BN_REQUIRE(Mt >= 2, "at least one inducing point (two transitions) is needed");
    BN_REQUIRE(mean && cov && gain && joint_mean && joint_cov, "null array");
    const unsigned grid = (unsigned)((Mt + 127) / 128);
#define X(FAM)                                                                                                    \
    if (k->family == FAM) {                                                                                       \
        MaternGen<FAM, 1> gen; gen.spec = *k; gen.dt = nullptr;                                                   \
        BN_LAUNCH("build_joint", (cudaStream_t)stream,                                                            \
                  build_joint_kernel<FAM><<<grid, 128, 0, (cudaStream_t)stream>>>(gen, Mt, mean, cov, gain, joint_mean, joint_cov)); \
    }
    SPARSE_FAMILIES(X)
#undef X
    BN_CUDA(cudaGetLastError());
    return 0;
}

static int sparse_launch(bool update, const bn_kernel_spec* k, int likelihood, const SparseArgs& a, int Q, const double* cub_x,
                         const double* cub_w, cudaStream_t s) {
    Cub1 cub;
    make_cub1((likelihood == BN_LIK_GAUSSIAN || likelihood == BN_LIK_POISSON_EXP) ? 0 : Q, cub_x, cub_w, cub);
    const long long Mt = a.Mz + 1;
    const unsigned grid = (unsigned)((Mt * 32 + 127) / 128);
    bool done = false;
#define X(FAM, LIK)                                                                                               \
    if (!done && k->family == FAM && likelihood == LIK) {                                                         \
        MaternGen<FAM, 1> gen; gen.spec = *k; gen.dt = nullptr;                                                   \
        if (update) BN_LAUNCH("sparse_site_update", s, sparse_site_kernel<FAM, LIK, true><<<grid, 128, 0, s>>>(gen, a, cub)); \
        else BN_LAUNCH("sparse_expected_density", s, sparse_site_kernel<FAM, LIK, false><<<grid, 128, 0, s>>>(gen, a, cub)); \
        done = true;                                                                                              \
    }
    SPARSE_LIKS(X, BN_MATERN12) SPARSE_LIKS(X, BN_MATERN32) SPARSE_LIKS(X, BN_MATERN52)
#undef X
    BN_REQUIRE(done, "likelihood %d has no single-latent statistics on the sparse Markov path", likelihood);
    BN_CUDA(cudaGetLastError());
    return 0;
}

extern "C" size_t bn_sparse_workspace_bytes(int64_t Mz) { return (size_t)(3 * (Mz + 1) + 64) * sizeof(double); }

extern "C" int bn_sparse_site_update(const bn_kernel_spec* k, int likelihood, double lik_param, int64_t N, int64_t Mz,
                                     const double* x, const double* y, const double* z, const int64_t* start,
                                     const double* post_mean, const double* post_cov, int Q, const double* cub_x,
                                     const double* cub_w, double lr, int ensure_psd, double* nat1, double* nat2,
                                     double* site_mean, double* site_cov, double* diffs, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    if (int rc = sparse_check_spec(k)) return rc;
    BN_REQUIRE(N >= 0 && Mz >= 1, "bad sizes N = %lld, Mz = %lld", (long long)N, (long long)Mz);
    BN_REQUIRE(x && y && z && start && post_mean && post_cov && nat1 && nat2 && site_mean && site_cov, "null array");
    BN_REQUIRE(likelihood == BN_LIK_GAUSSIAN || likelihood == BN_LIK_POISSON_EXP || (Q >= 1 && Q <= kMaxQ1 && cub_x && cub_w),
               "a 1-D cubature rule (host arrays, Q <= %d) is needed", kMaxQ1);
    BN_REQUIRE(workspace && workspace_bytes >= bn_sparse_workspace_bytes(Mz), "workspace too small");
    SparseArgs a{N, Mz, x, y, z, (const long long*)start, post_mean, post_cov, lik_param, lr, ensure_psd, nat1, nat2,
                 site_mean, site_cov, (double*)workspace};
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = sparse_launch(true, k, likelihood, a, Q, cub_x, cub_w, s)) return rc;
    if (diffs) {
        const long long Mt = Mz + 1;
        const int p = 2 * family_dim(k->family);
        BN_LAUNCH("sum", s, sum_kernel<false><<<1, 1024, 0, s>>>(a.partials, Mt, diffs, 1.0 / (double)(Mt * p)));
        BN_LAUNCH("sum", s, sum_kernel<false><<<1, 1024, 0, s>>>(a.partials + Mt, Mt, diffs + 1, 1.0 / (double)(Mt * p * p)));
        BN_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int bn_sparse_expected_density(const bn_kernel_spec* k, int likelihood, double lik_param, int64_t N, int64_t Mz,
                                          const double* x, const double* y, const double* z, const int64_t* start,
                                          const double* post_mean, const double* post_cov, int Q, const double* cub_x,
                                          const double* cub_w, double* sum, void* workspace, size_t workspace_bytes,
                                          void* stream) {
    if (int rc = sparse_check_spec(k)) return rc;
    BN_REQUIRE(N >= 0 && Mz >= 1, "bad sizes N = %lld, Mz = %lld", (long long)N, (long long)Mz);
    BN_REQUIRE(x && y && z && start && post_mean && post_cov && sum, "null array");
    BN_REQUIRE(likelihood == BN_LIK_GAUSSIAN || likelihood == BN_LIK_POISSON_EXP || (Q >= 1 && Q <= kMaxQ1 && cub_x && cub_w),
               "a 1-D cubature rule (host arrays, Q <= %d) is needed", kMaxQ1);
    BN_REQUIRE(workspace && workspace_bytes >= bn_sparse_workspace_bytes(Mz), "workspace too small");
    SparseArgs a{N, Mz, x, y, z, (const long long*)start, post_mean, post_cov, lik_param, 1.0, 0, nullptr, nullptr, nullptr,
                 nullptr, (double*)workspace};
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = sparse_launch(false, k, likelihood, a, Q, cub_x, cub_w, s)) return rc;
    BN_LAUNCH("sum", s, sum_kernel<false><<<1, 1024, 0, s>>>(a.partials + 2 * (Mz + 1), Mz + 1, sum, 1.0));
    BN_CUDA(cudaGetLastError());
    return 0;
}
