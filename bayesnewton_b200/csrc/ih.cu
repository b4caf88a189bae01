// Infinite-horizon (steady-state) Kalman filter / RTS smoother: C ABI bn_ih_filter, bn_ih_smoother.
//
// Reference: kalman_filter_infinite_horizon (ops.py:881-952) with _sequential_kf_ih / _parallel_kf_ih (:827-878) and
// rauch_tung_striebel_smoother_infinite_horizon (:1018-1068) with _sequential_rts_ih / _parallel_rts_ih (:978-1015).
// Once the state covariance is frozen at the Riccati fixed point Pdare (found on the host: 20 iterations of d x d
// algebra, ops.py:796-824), both passes are AFFINE recursions in the mean,
//     filter    m_k  = (A - K_k H A) m_{k-1} + K_k y_k,     K_k = Pdare H^T / (H Pdare H^T + R_k)
//     smoother  sm_k = G sm_{k+1} + (I - G A) fm_k,          G = cov A^T Pdare^-1
// so the temporally parallel form is a scan over (M, v) pairs: one thread per chunk of steps composes its maps
// (phase 1), the chunk maps are scanned (scan.cuh, the same multi-level kernel as the full filter), and each thread
// re-runs its chunk from its incoming mean (phase 3), forming M_k from (y_k, R_k) in registers -- 16 B in, 8 d B out
// per step for the filter.  The reference's heteroscedastic scan composes the INVERSES of the contractions
// (ops.py:849-853, 930-933) and overflows on long series; the maps themselves are composed here.
// H = e_0^T (one latent, one site per step), d <= 4.
#include "common.cuh"
#include "core.cuh"
#include "scan.cuh"

namespace bn {

template <int d>
struct AffineAlg {
    static constexpr int kElem = d * d + d;
    static constexpr int kState = d;
    struct Elem { double M[d * d], v[d]; };
    struct State { double m[d]; };
    static BN_DEV void identity(Elem& e) {
#pragma unroll
        for (int i = 0; i < d; ++i) {
            e.v[i] = 0.0;
#pragma unroll
            for (int j = 0; j < d; ++j) e.M[i * d + j] = (i == j) ? 1.0 : 0.0;
        }
    }
    // e1 first, then e2:  x -> M2 (M1 x + v1) + v2
    static BN_DEV void combine(const Elem& e1, const Elem& e2, Elem& out) {
#pragma unroll
        for (int i = 0; i < d; ++i) {
            double s = e2.v[i];
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(e2.M[i * d + l], e1.v[l], s);
            out.v[i] = s;
#pragma unroll
            for (int j = 0; j < d; ++j) {
                double t = 0.0;
#pragma unroll
                for (int l = 0; l < d; ++l) t = fma(e2.M[i * d + l], e1.M[l * d + j], t);
                out.M[i * d + j] = t;
            }
        }
    }
    static BN_DEV void apply(const Elem& e, const State& s, State& out) {
#pragma unroll
        for (int i = 0; i < d; ++i) {
            double t = e.v[i];
#pragma unroll
            for (int l = 0; l < d; ++l) t = fma(e.M[i * d + l], s.m[l], t);
            out.m[i] = t;
        }
    }
    static BN_DEV void load(const double* base, long long stride, long long i, Elem& e) {
        const double* p = base + i;
#pragma unroll
        for (int k = 0; k < d * d; ++k) e.M[k] = p[k * stride];
#pragma unroll
        for (int k = 0; k < d; ++k) e.v[k] = p[(d * d + k) * stride];
    }
    static BN_DEV void store(double* base, long long stride, long long i, const Elem& e) {
        double* p = base + i;
#pragma unroll
        for (int k = 0; k < d * d; ++k) p[k * stride] = e.M[k];
#pragma unroll
        for (int k = 0; k < d; ++k) p[(d * d + k) * stride] = e.v[k];
    }
    static __device__ __forceinline__ void shfl_up(Elem& e, int delta) {
        double* p = reinterpret_cast<double*>(&e);
#pragma unroll
        for (int k = 0; k < kElem; ++k) p[k] = __shfl_up_sync(0xffffffffu, p[k], delta);
    }
};

// ---- the two recursions as generators of (M_k, v_k)
template <int d>
struct IhFilterGen {
    double A[d * d], p[d], s0;  // transition, Pdare H^T, H Pdare H^T
    const double* y;
    const double* R;
    long long r_stride;         // 1: per-step pseudo variances (heteroscedastic); 0: the tied value for every step
    BN_DEV void elem(long long k, typename AffineAlg<d>::Elem& e, double& S) const {
        S = s0 + R[k * r_stride];
        const double L = sqrt(S), inv = (1.0 / L) / L;  // solve(S, H) by Cholesky, utils.py:14-19
        const double yk = y[k];
#pragma unroll
        for (int i = 0; i < d; ++i) {
            const double kk = p[i] * inv;
            e.v[i] = kk * yk;
#pragma unroll
            for (int j = 0; j < d; ++j) e.M[i * d + j] = fma(-kk, A[j], A[i * d + j]);  // A - K (H A), H A = row 0 of A
        }
    }
};

template <int d>
struct IhSmootherGen {
    double G[d * d], IGA[d * d];  // gain, I - G A
    const double* fm;             // [N, d]
    BN_DEV void elem(long long k, typename AffineAlg<d>::Elem& e, double& S) const {
        S = 0.0;
#pragma unroll
        for (int i = 0; i < d; ++i) {
            double s = 0.0;
#pragma unroll
            for (int l = 0; l < d; ++l) s = fma(IGA[i * d + l], fm[k * d + l], s);
            e.v[i] = s;
#pragma unroll
            for (int j = 0; j < d; ++j) e.M[i * d + j] = G[i * d + j];
        }
    }
};

// phase 1: composition of the chunk's maps.  REVERSE: the recursion runs from the last step down (smoother); the chunk
// that is processed first (scan position 0) is then the LAST one
template <int d, class Gen, bool REVERSE>
__global__ void __launch_bounds__(kChunkThreads)
ih_reduce_kernel(Gen gen, long long N, int L, long long nchunks, double* agg) {
    using Alg = AffineAlg<d>;
    const long long c = (long long)blockIdx.x * kChunkThreads + threadIdx.x;
    if (c >= nchunks) return;
    const long long k0 = c * L, k1 = (k0 + L < N) ? k0 + L : N;
    typename Alg::Elem acc, e, r;
    Alg::identity(acc);
    double S;
    if (!REVERSE) {
        for (long long k = k0; k < k1; ++k) { gen.elem(k, e, S); Alg::combine(acc, e, r); acc = r; }
        Alg::store(agg, nchunks, c, acc);
    } else {
        for (long long k = k1 - 1; k >= k0; --k) { gen.elem(k, e, S); Alg::combine(acc, e, r); acc = r; }
        Alg::store(agg, nchunks, nchunks - 1 - c, acc);
    }
}

// phase 3 (and the whole sequential form when nchunks == 1): the recursion from the chunk's incoming mean.
// Filter: means[N, d], log-likelihood partial per chunk (mvn_logpdf with the mask rule, utils.py:376-396);
// smoother: out[N] = H sm or out[N, d] = sm.
template <int d, class Gen, bool REVERSE>
__global__ void __launch_bounds__(kChunkThreads)
ih_apply_kernel(Gen gen, long long N, int L, long long nchunks, const double* prefix, const double* init,
                const double* hrow, const unsigned char* mask, int return_full, double* out, double* ell_partials) {
    using Alg = AffineAlg<d>;
    const long long c = (long long)blockIdx.x * kChunkThreads + threadIdx.x;
    if (c >= nchunks) return;
    const long long k0 = c * L, k1 = (k0 + L < N) ? k0 + L : N;
    typename Alg::State s, t;
#pragma unroll
    for (int i = 0; i < d; ++i) s.m[i] = init[i];
    const long long pos = REVERSE ? nchunks - 1 - c : c;
    if (pos > 0) {
        typename Alg::Elem pe;
        Alg::load(prefix, nchunks, pos - 1, pe);
        Alg::apply(pe, s, t);
        s = t;
    }
    typename Alg::Elem e;
    double S, ell = 0.0;
    if constexpr (!REVERSE) {
        for (long long k = k0; k < k1; ++k) {
            gen.elem(k, e, S);
            if (ell_partials) {
                double om = 0.0;  // H A m_{k-1}
#pragma unroll
                for (int l = 0; l < d; ++l) om = fma(hrow[l], s.m[l], om);
                if (!(mask && mask[k])) {
                    const double Lc = sqrt(S), diff = gen.y[k] - om;
                    ell += -0.5 * (diff * ((diff / Lc) / Lc) + kLog2Pi + 2.0 * log(fabs(Lc)));
                }
            }
            Alg::apply(e, s, t);
            s = t;
#pragma unroll
            for (int i = 0; i < d; ++i) out[k * d + i] = s.m[i];
        }
        if (ell_partials) ell_partials[c] = ell;
    } else {
        for (long long k = k1 - 1; k >= k0; --k) {
            gen.elem(k, e, S);
            Alg::apply(e, s, t);
            s = t;
            if (return_full) {
#pragma unroll
                for (int i = 0; i < d; ++i) out[k * d + i] = s.m[i];
            } else {
                out[k] = s.m[0];
            }
        }
    }
}

template <int d>
static size_t ih_ws_doubles(long long N) {
    ChunkPlan cp = plan_chunks(N > 0 ? N : 1);
    return 64 + scan_plan_doubles(cp.nchunks, AffineAlg<d>::kElem) + cp.nchunks;
}

template <int d, class Gen, bool REVERSE>
static int ih_run(const Gen& gen, int form, long long N, const double* init_dev, const double* hrow_host,
                  const unsigned char* mask, int return_full, double* out, double* ell, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
    using Alg = AffineAlg<d>;
    const size_t need = ih_ws_doubles<d>(N) * sizeof(double);
    BN_REQUIRE(ws != nullptr && ws_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    double* p = (double*)ws;
    double* dconst = p; p += 64;  // [0, d): zero initial mean, [d, 2d): H A
    double hostc[64] = {0};
    for (int i = 0; i < d; ++i) hostc[d + i] = hrow_host ? hrow_host[i] : 0.0;
    BN_CUDA(cudaMemcpyAsync(dconst, hostc, sizeof(double) * 2 * d, cudaMemcpyHostToDevice, st));
    const double* init = init_dev ? init_dev : dconst;
    ChunkPlan cp = plan_chunks(N);
    if (form == BN_SEQUENTIAL) cp.nchunks = 1;  // one thread walks the whole series in time order
    ScanPlan plan = make_scan_plan(p, cp.nchunks, Alg::kElem);
    p += scan_plan_doubles(cp.nchunks, Alg::kElem);
    double* partials = p;
    const unsigned grid = (unsigned)((cp.nchunks + kChunkThreads - 1) / kChunkThreads);
    const int L = (form == BN_SEQUENTIAL) ? (int)(N < 2147483647LL ? N : 2147483647LL) : cp.L;
    BN_REQUIRE(form != BN_SEQUENTIAL || N < 2147483647LL, "sequential form: N too large");
    if (cp.nchunks > 1) {
        BN_LAUNCH("ih_reduce", st, (ih_reduce_kernel<d, Gen, REVERSE><<<grid, kChunkThreads, 0, st>>>(gen, N, L, cp.nchunks, plan.input0)));
        BN_CUDA(cudaGetLastError());
        BN_CUDA(run_scan<Alg>(plan, st));
    }
    BN_LAUNCH("ih_apply", st, (ih_apply_kernel<d, Gen, REVERSE><<<grid, kChunkThreads, 0, st>>>(
                                  gen, N, L, cp.nchunks, plan.prefix[0], init, dconst + d, mask, return_full, out,
                                  ell ? partials : nullptr)));
    BN_CUDA(cudaGetLastError());
    if (ell) {
        sum_kernel<false><<<1, 1024, 0, st>>>(partials, cp.nchunks, ell, 1.0);
        BN_CUDA(cudaGetLastError());
    }
    return 0;
}

template <int d>
static int ih_filter_d(int form, long long N, const double* A, const double* Pdare, const double* y, const double* R,
                       int r_is_scalar, const unsigned char* mask, double* ell, double* means, void* ws, size_t nb,
                       cudaStream_t st) {
    IhFilterGen<d> g;
    for (int i = 0; i < d * d; ++i) g.A[i] = A[i];
    for (int i = 0; i < d; ++i) g.p[i] = Pdare[i * d];
    g.s0 = Pdare[0];
    g.y = y;
    g.R = R;
    g.r_stride = r_is_scalar ? 0 : 1;
    return ih_run<d, IhFilterGen<d>, false>(g, form, N, nullptr, A /* row 0 = H A */, mask, 1, means, ell, ws, nb, st);
}

template <int d>
static int ih_smoother_d(int form, long long N, const double* A, const double* gain, const double* fm,
                         int return_full, double* out, void* ws, size_t nb, cudaStream_t st) {
    IhSmootherGen<d> g;
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int l = 0; l < d; ++l) s -= gain[i * d + l] * A[l * d + j];
            g.IGA[i * d + j] = s;
            g.G[i * d + j] = gain[i * d + j];
        }
    g.fm = fm;
    // the carry starts at the last filtered mean (ops.py:988-990) and the last step is processed like every other
    return ih_run<d, IhSmootherGen<d>, true>(g, form, N, fm + (N - 1) * d, nullptr, nullptr, return_full, out, nullptr, ws, nb, st);
}
}  // namespace bn

using namespace bn;

extern "C" size_t bn_ih_workspace_bytes(int d, int64_t N) {
    if (d < 1 || d > 4 || N < 0) return 0;
    size_t n = 0;
    switch (d) {
        case 1: n = ih_ws_doubles<1>(N); break;
        case 2: n = ih_ws_doubles<2>(N); break;
        case 3: n = ih_ws_doubles<3>(N); break;
        case 4: n = ih_ws_doubles<4>(N); break;
    }
    return (n + 64) * sizeof(double);
}

extern "C" int bn_ih_filter(int form, int d, int64_t N, const double* A_host, const double* Pdare_host, const double* y,
                            const double* noise_var, int noise_is_scalar, const uint8_t* mask, double* ell, double* means,
                            void* workspace, size_t workspace_bytes, void* stream) {
    BN_REQUIRE(d >= 1 && d <= 4, "state dimension %d outside 1..4", d);
    BN_REQUIRE(N > 0, "N must be positive");
    BN_REQUIRE(A_host && Pdare_host && y && noise_var && means, "null array");
    BN_REQUIRE(form == BN_SEQUENTIAL || form == BN_SCAN, "unknown form %d", form);
    cudaStream_t st = (cudaStream_t)stream;
    switch (d) {
        case 1: return ih_filter_d<1>(form, N, A_host, Pdare_host, y, noise_var, noise_is_scalar, mask, ell, means, workspace, workspace_bytes, st);
        case 2: return ih_filter_d<2>(form, N, A_host, Pdare_host, y, noise_var, noise_is_scalar, mask, ell, means, workspace, workspace_bytes, st);
        case 3: return ih_filter_d<3>(form, N, A_host, Pdare_host, y, noise_var, noise_is_scalar, mask, ell, means, workspace, workspace_bytes, st);
        default: return ih_filter_d<4>(form, N, A_host, Pdare_host, y, noise_var, noise_is_scalar, mask, ell, means, workspace, workspace_bytes, st);
    }
}

extern "C" int bn_ih_smoother(int form, int d, int64_t N, const double* A_host, const double* gain_host,
                              const double* filter_mean, int return_full, double* means,
                              void* workspace, size_t workspace_bytes, void* stream) {
    BN_REQUIRE(d >= 1 && d <= 4, "state dimension %d outside 1..4", d);
    BN_REQUIRE(N > 0, "N must be positive");
    BN_REQUIRE(A_host && gain_host && filter_mean && means, "null array");
    BN_REQUIRE(form == BN_SEQUENTIAL || form == BN_SCAN, "unknown form %d", form);
    cudaStream_t st = (cudaStream_t)stream;
    switch (d) {
        case 1: return ih_smoother_d<1>(form, N, A_host, gain_host, filter_mean, return_full, means, workspace, workspace_bytes, st);
        case 2: return ih_smoother_d<2>(form, N, A_host, gain_host, filter_mean, return_full, means, workspace, workspace_bytes, st);
        case 3: return ih_smoother_d<3>(form, N, A_host, gain_host, filter_mean, return_full, means, workspace, workspace_bytes, st);
        default: return ih_smoother_d<4>(form, N, A_host, gain_host, filter_mean, return_full, means, workspace, workspace_bytes, st);
    }
}
