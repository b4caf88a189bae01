// Prediction at test inputs: the step after inference in every demo.
//   bn_temporal_conditional   temporal_conditional -> predict_from_state -> compute_conditional_statistics
//                             (bayesnewton/utils.py:99-136, 173-215) + the H projection of predict()
//                             (basemodels.py:766-816): one thread per test point, everything in registers,
//                             A(dt), Q(dt) of both neighbouring gaps generated in place from the closed forms.
//   bn_likelihood_predict     Likelihood.predict / predict_cubature (likelihoods.py:493-506, 802-803;
//                             cubature.py:438-465): E[y], Var[y] at each test point.
#include "common.cuh"
#include "gen.cuh"
#include "condstats.cuh"

namespace bn {

struct TcArgs {
    long long N, Ns;
    const double* x;       // [N] sorted training inputs
    const double* xs;      // [Ns] test inputs
    const double* mean;    // [N,d]   smoothed state means
    const double* cov;     // [N,d,d] smoothed state covariances
    const double* gain;    // [N,d,d] smoother gains
    int return_full;
    double* out_mean;      // [Ns,Df] or [Ns,d]
    double* out_cov;       // [Ns,Df,Df] or [Ns,d,d]
};

template <class Gen>
__global__ void __launch_bounds__(128) temporal_conditional_kernel(Gen gen, TcArgs a) {
    constexpr int d = Gen::d, D = Gen::D, n = Gen::n;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.Ns) return;
    const double xt = a.xs[t];
    // ind = searchsorted(X_aug, xt) - 1 with X_aug = [-1e10, x, 1e10] (basemodels.py:793-794, utils.py:131)
    long long lo = 0, hi = a.N;  // number of training inputs strictly below xt
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (a.x[mid] < xt) lo = mid + 1; else hi = mid;
    }
    long long ind = lo;                       // index into the augmented arrays
    if (xt > 1e10) ind += 1;                  // beyond the dummy state (degenerate, kept for fidelity)
    if (!(xt > -1e10)) ind = -1;
    if (ind < 0 || ind > a.N) {               // outside the dummy states: the reference would index out of range
        const int od = a.return_full ? d : D;
        for (int i = 0; i < od; ++i) a.out_mean[t * od + i] = nan("");
        for (int i = 0; i < od * od; ++i) a.out_cov[t * od * od + i] = nan("");
        return;
    }
    const double xl = ind == 0 ? -1e10 : a.x[ind - 1];
    const double xr = ind == a.N ? 1e10 : a.x[ind];
    double P1[d * d], W[d * d], Tm[d * d];
    cond_stats(gen, xt - xl, xr - xt, P1, W, Tm);
    // neighbouring smoothed states: augmented with (minf, Pinf) at both ends, gains with a leading zero (utils.py:124-128)
    double ml[d], mr[d], Cl[d * d], Cr[d * d], G[d * d];
    double Pinf[symn(d)];
    gen.pinf(Pinf);
#pragma unroll
    for (int i = 0; i < d; ++i) {
        ml[i] = ind == 0 ? 0.0 : a.mean[(ind - 1) * d + i];
        mr[i] = ind == a.N ? 0.0 : a.mean[ind * d + i];
#pragma unroll
        for (int j = 0; j < d; ++j) {
            Cl[i * d + j] = ind == 0 ? Pinf[sidx(i, j)] : a.cov[(ind - 1) * d * d + i * d + j];
            Cr[i * d + j] = ind == a.N ? Pinf[sidx(i, j)] : a.cov[ind * d * d + i * d + j];
            G[i * d + j] = ind == 0 ? 0.0 : a.gain[(ind - 1) * d * d + i * d + j];
        }
    }
    // mean = P [ml; mr];  cov = P [[Cl, X],[X^T, Cr]] P^T + T with X = G Cr   (utils.py:113-120)
    double X[d * d];
    matmul<d, d, d>(G, Cr, X);
    double m[d];
#pragma unroll
    for (int i = 0; i < d; ++i) {
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < d; ++l) s = fma(P1[i * d + l], ml[l], fma(W[i * d + l], mr[l], s));
        m[i] = s;
    }
    // U1 = P1 Cl + W X^T,  U2 = P1 X + W Cr;  cov = U1 P1^T + U2 W^T + T
    double U1[d * d], U2[d * d];
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int l = 0; l < d; ++l) {
                s1 = fma(P1[i * d + l], Cl[l * d + j], fma(W[i * d + l], X[j * d + l], s1));
                s2 = fma(P1[i * d + l], X[l * d + j], fma(W[i * d + l], Cr[l * d + j], s2));
            }
            U1[i * d + j] = s1;
            U2[i * d + j] = s2;
        }
    double C[d * d];
#pragma unroll
    for (int i = 0; i < d; ++i)
#pragma unroll
        for (int j = 0; j < d; ++j) {
            double s = Tm[i * d + j];
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(U1[i * d + l], P1[j * d + l], fma(U2[i * d + l], W[j * d + l], s));
            C[i * d + j] = s;
        }
    if (a.return_full) {
#pragma unroll
        for (int i = 0; i < d; ++i) a.out_mean[t * d + i] = m[i];
#pragma unroll
        for (int i = 0; i < d * d; ++i) a.out_cov[t * d * d + i] = C[i];
    } else {  // H = blockdiag([1 0 ..]) selects the first state of each component
#pragma unroll
        for (int i = 0; i < D; ++i) {
            a.out_mean[t * D + i] = m[i * n];
#pragma unroll
            for (int j = 0; j < D; ++j) a.out_cov[t * D * D + i * D + j] = C[(i * n) * d + j * n];
        }
    }
}

// E[y], Var[y] for a scalar latent per point: Gaussian closed form, Bernoulli by 1-D cubature
__device__ __forceinline__ double probit_p(double f) { return 0.5 * (1.0 + erf(f * 0.7071067811865476)) * (1.0 - 2e-3) + 1e-3; }

__global__ void likelihood_predict_kernel(int lik, double param, double param2, long long N, const double* mean_f, const double* var_f,
                                          int Q, const double* cx, const double* cw, double* mean_y, double* var_y) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const double m = mean_f[t], v = var_f[t];
    if (lik == BN_LIK_GAUSSIAN) {
        mean_y[t] = m;
        var_y[t] = v + param;
        return;
    }
    const double sd = sqrt(v);  // 1 x 1 Cholesky factor (NaN for a negative variance, as cho_factor)
    double e1 = 0.0, e2 = 0.0;
    for (int i = 0; i < Q; ++i) {
        const double f = fma(sd, cx[i], m);
        if (lik == BN_LIK_POISSON_EXP) {  // E[y|f] = Var[y|f] = b exp(f)  (likelihoods.py:952-959)
            const double mu = param * exp(f);
            e1 = fma(cw[i], mu, e1);
            e2 = fma(cw[i], mu + mu * mu, e2);
            continue;
        }
        if (lik == BN_LIK_STUDENTS_T || lik == BN_LIK_GAMMA_EXP || lik == BN_LIK_NEGBIN_EXP || lik == BN_LIK_BETA_PROBIT) {
            double E, V;  // conditional moments, likelihoods.py:1043-1044, 1095-1097, 1136-1138, 1184-1189
            if (lik == BN_LIK_STUDENTS_T) {
                E = f; V = (param * param) * (param2 / (param2 - 2.0));
            } else if (lik == BN_LIK_GAMMA_EXP) {
                const double sc = exp(f);
                E = param * sc; V = param * (sc * sc);
            } else if (lik == BN_LIK_NEGBIN_EXP) {
                E = exp(f) * param2; V = E + E * E * param;
            } else {
                E = probit_p(f); V = (E - E * E) / (param + 1.0);
            }
            e1 = fma(cw[i], E, e1);
            e2 = fma(cw[i], V + E * E, e2);
            continue;
        }
        const double p = lik == BN_LIK_BERNOULLI_PROBIT ? probit_p(f) : 1.0 / (1.0 + exp(-f));
        e1 = fma(cw[i], p, e1);
        e2 = fma(cw[i], p * (1.0 - p) + p * p, e2);  // Cov[y|f] + E[y|f]^2  (likelihoods.py:854-860)
    }
    mean_y[t] = e1;
    var_y[t] = e2 - e1 * e1;
}

}  // namespace bn

using namespace bn;

extern "C" int bn_temporal_conditional(const bn_kernel_spec* k, int64_t N, const double* x, int64_t Ns, const double* x_test,
                                       const double* mean, const double* cov, const double* gain, int return_full,
                                       double* out_mean, double* out_cov, void* stream) {
    BN_REQUIRE(k != nullptr, "kernel spec is null");
    BN_REQUIRE(N >= 1 && Ns >= 0, "bad sizes N = %lld, N_test = %lld", (long long)N, (long long)Ns);
    if (Ns == 0) return 0;
    BN_REQUIRE(x && x_test && mean && cov && gain && out_mean && out_cov, "null array");
    TcArgs a{N, Ns, x, x_test, mean, cov, gain, return_full, out_mean, out_cov};
    const unsigned grid = (unsigned)((Ns + 127) / 128);
#define X(FAM, NC)                                                                                   \
    if (k->family == FAM && k->n_components == NC) {                                                 \
        MaternGen<FAM, NC> gen;                                                                      \
        gen.spec = *k;                                                                               \
        gen.dt = nullptr;                                                                            \
        BN_LAUNCH("temporal_conditional", (cudaStream_t)stream,                                      \
                  temporal_conditional_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(gen, a));      \
        BN_CUDA(cudaGetLastError());                                                                 \
        return 0;                                                                                    \
    }
    BN_GROUP_M_A(X) X(BN_MATERN32, 2) X(BN_MATERN52, 1) BN_GROUP_M_D(X)  // state dimension <= 4
#undef X
    set_error("unsupported kernel spec for prediction: family %d with %d components", k->family, k->n_components);
    return -1;
}

extern "C" int bn_likelihood_predict(int likelihood, double lik_param, int64_t N, const double* mean_f, const double* var_f,
                                     int Q, const double* cub_x, const double* cub_w, double* mean_y, double* var_y,
                                     void* stream) {
    return bn_likelihood_predict2(likelihood, lik_param, 0.0, N, mean_f, var_f, Q, cub_x, cub_w, mean_y, var_y, stream);
}

extern "C" int bn_likelihood_predict2(int likelihood, double lik_param, double lik_param2, int64_t N, const double* mean_f,
                                      const double* var_f, int Q, const double* cub_x, const double* cub_w, double* mean_y,
                                      double* var_y, void* stream) {
    BN_REQUIRE(likelihood == BN_LIK_GAUSSIAN || likelihood == BN_LIK_BERNOULLI_PROBIT || likelihood == BN_LIK_BERNOULLI_LOGIT ||
                   likelihood == BN_LIK_POISSON_EXP || likelihood == BN_LIK_STUDENTS_T || likelihood == BN_LIK_GAMMA_EXP ||
                   likelihood == BN_LIK_NEGBIN_EXP || likelihood == BN_LIK_BETA_PROBIT,
               "likelihood %d has no single-latent predict on this path", likelihood);
    BN_REQUIRE(N >= 0, "N must be non-negative");
    if (N == 0) return 0;
    BN_REQUIRE(mean_f && var_f && mean_y && var_y, "null array");
    BN_REQUIRE(likelihood == BN_LIK_GAUSSIAN || (Q >= 1 && cub_x && cub_w), "a cubature rule (device arrays) is needed");
    const unsigned grid = (unsigned)((N + 255) / 256);
    BN_LAUNCH("likelihood_predict", (cudaStream_t)stream,
              likelihood_predict_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(likelihood, lik_param, lik_param2, N, mean_f, var_f, Q, cub_x,
                                                                             cub_w, mean_y, var_y));
    BN_CUDA(cudaGetLastError());
    return 0;
}
